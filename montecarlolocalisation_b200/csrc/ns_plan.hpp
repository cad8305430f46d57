// ns_plan.hpp — host-side plan of the globally exact systematic resampling across shards (pure integer arithmetic;
// no GPU needed, so the multi-rank protocol is testable on CPU with gloo).
//
// Shard g holds a contiguous range of the global particle order with weight total T_g; O_g = sum_{h<g} T_h. Output slot
// k selects ancestor(k) = min{ i : C_i * (N<<32) > ((k<<32)+u0) * T } (ns_core.cuh). The slots whose ancestor lives in
// shard g form the contiguous range [first_slot(O_g), first_slot(O_g + T_g)).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NSP_HD __host__ __device__
#else
#define NSP_HD
#endif

namespace mcl {
namespace ns {

// min{ k in [0,N] : ((k<<32)+u0) * T >= O * (N<<32) }   (T > 0, O <= T)
NSP_HD inline int64_t first_slot(uint64_t O, uint64_t T, uint64_t N, uint32_t u0) {
    if (O == 0) return 0;
    if (O >= T) return (int64_t)N;
    const unsigned __int128 rhs = ((unsigned __int128)O * N) << 32;           // < 2^60 * 2^31 * 2^32
    const unsigned __int128 x_min = (rhs + T - 1) / T;                        // smallest X with X*T >= rhs
    if (x_min <= u0) return 0;
    const unsigned __int128 k = (x_min - u0 + (((unsigned __int128)1 << 32) - 1)) >> 32;
    return k > N ? (int64_t)N : (int64_t)k;
}

// contiguous split of n_global particles over `world` shards: every shard holds per = ceil(n_global/world) slots except
// the last ones, which hold what is left; slot k lives on shard k / per.
inline void shard_range(int64_t n_global, int world, int rank, int64_t* begin, int64_t* count, int64_t* per_rank) {
    const int64_t per = (n_global + world - 1) / world;
    int64_t b = (int64_t)rank * per;
    if (b > n_global) b = n_global;
    int64_t c = n_global - b;
    if (c > per) c = per;
    *begin = b; *count = c; *per_rank = per;
}

}  // namespace ns
}  // namespace mcl
