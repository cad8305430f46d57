// exact_scan_model.cpp — CPU model check of montecarlolocalisation_b200/csrc/exact_scan_core.cuh.
// Runs the tile / classify / parity-monoid / chain / apply pipeline exactly as the CUDA kernels structure it (but
// serially), on adversarial random weight vectors, and compares every prefix value bit-for-bit with the plain
// left-to-right f64 loop it must reproduce. Exit code 0 = all cases identical.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../montecarlolocalisation_b200/csrc/exact_scan_core.cuh"

using namespace mcl::xs;

static const int TILE = 2048;
static const int SEQ_CAP = 16;

struct SeqEntry { uint32_t idx; float w; Par pre; int first_in_tile; int E_prev; };
struct Stats { long seq = 0, fallback = 0, cases = 0, uniform_tiles = 0; };

// returns false => the algorithm asked for the sequential fallback
static bool exact_scan(const std::vector<float>& w, std::vector<double>& out, int order, Stats& st) {
    const size_t n = w.size();
    const size_t nt = (n + TILE - 1) / TILE;
    out.assign(n, 0.0);
    // pass 1+2: tile sums in some non-sequential association, then offsets
    std::vector<double> toff(nt + 1, 0.0), tsum_of(nt, 0.0);
    std::vector<double> ptilde(n);
    for (size_t t = 0; t < nt; t++) {
        size_t a = t * TILE, b = std::min(n, a + (size_t)TILE);
        double s = 0;
        if (order == 0) { for (size_t i = a; i < b; i++) s += w[i]; }
        else {   // 8-wide thread-serial then pairwise, like the block scan
            std::vector<double> part;
            for (size_t i = a; i < b; i += 8) { double p = 0; for (size_t j = i; j < std::min(b, i + 8); j++) p += w[j]; part.push_back(p); }
            while (part.size() > 1) { std::vector<double> nx; for (size_t i = 0; i + 1 < part.size(); i += 2) nx.push_back(part[i] + part[i + 1]); if (part.size() & 1) nx.push_back(part.back()); part.swap(nx); }
            s = part.empty() ? 0 : part[0];
        }
        toff[t + 1] = toff[t] + s;
        tsum_of[t] = s;
    }
    if (order == 2) {
        // the one-kernel form (exact_scan_fused.cuh): the value at every tile edge is F(t) = a fixed-association sum of the lower
        // tile sums (256 lane-strided partials in ascending order, then a tree), evaluated independently for every t
        for (size_t t = 0; t <= nt; t++) {
            double part[256] = {0};
            for (size_t k = 0; k < t; k++) part[k % 256] += tsum_of[k];
            for (int stride = 128; stride > 0; stride >>= 1) for (int l = 0; l < stride; l++) part[l] += part[l + stride];
            toff[t] = part[0];
        }
    }
    for (size_t t = 0; t < nt; t++) {
        size_t a = t * TILE, b = std::min(n, a + (size_t)TILE);
        double run = 0;
        for (size_t i = a; i < b; i++) { run += w[i]; ptilde[i] = toff[t] + run; }
        ptilde[b - 1] = toff[t + 1];            // tile edges are shared values
    }
    const uint64_t depth = nt + 32;
    // pass 3: classify + in-tile segmented composites + SEQ entries
    std::vector<Par> vlast(nt);
    std::vector<int> has_seq(nt, 0), seq_count(nt, 0);
    std::vector<std::vector<SeqEntry>> entries(nt);
    std::vector<uint8_t> kind(n);              // 0 PAR, 1 SEQ
    std::vector<Par> V(n);
    std::vector<int> Eof(n);
    for (size_t t = 0; t < nt; t++) {
        size_t a = t * TILE, b = std::min(n, a + (size_t)TILE);
        Par v = par_identity();
        // one-kernel form: a tile whose two edge values lie safely inside one binade is uniform - every element is PAR in that
        // binade and the per-element predictions are skipped
        const Pred e0 = predict(toff[t], margin_for(b, depth)), e1 = predict(toff[t + 1], margin_for(b, depth));
        const bool uniform = order == 2 && t > 0 && e0.ok && e1.ok && e0.E == e1.E;
        if (uniform) st.uniform_tiles++;
        for (size_t i = a; i < b; i++) {
            Pred cur = predict(ptilde[i], margin_for(i, depth));
            Pred prev = predict(i ? ptilde[i - 1] : 0.0, margin_for(i ? i - 1 : 0, depth));
            if (uniform) { cur.ok = prev.ok = true; cur.zero = false; cur.E = prev.E = e0.E; }
            Par f;
            bool par;
            if (cur.zero) { par = true; f = par_identity(); }
            else {
                par = cur.ok && prev.ok && cur.E == prev.E;
                if (par && !par_of_weight(w[i], cur.E, f)) return false;
            }
            Eof[i] = cur.E;
            if (par) { v = par_compose(v, f); kind[i] = 0; V[i] = v; }
            else {
                kind[i] = 1;
                if ((int)entries[t].size() >= SEQ_CAP) return false;
                entries[t].push_back(SeqEntry{(uint32_t)i, w[i], v, entries[t].empty() ? 1 : 0, prev.E});
                v = par_identity();
                V[i] = v;
                has_seq[t] = 1;
            }
        }
        vlast[t] = v;
        seq_count[t] = (int)entries[t].size();
        st.seq += seq_count[t];
    }
    // pass 4: chain
    std::vector<Par> carry(nt + 1);
    std::vector<int> seq_base(nt + 1, 0);
    carry[0] = par_identity();
    for (size_t t = 0; t < nt; t++) {
        carry[t + 1] = has_seq[t] ? vlast[t] : par_compose(carry[t], vlast[t]);
        seq_base[t + 1] = seq_base[t] + seq_count[t];
    }
    std::vector<double> seq_s(seq_base[nt]);
    double s = 0.0;
    bool ok = true;
    for (size_t t = 0; t < nt; t++)
        for (int k = 0; k < seq_count[t]; k++) {
            const SeqEntry& e = entries[t][k];
            Par comp = e.first_in_tile ? par_compose(carry[t], e.pre) : e.pre;
            s = par_apply(s, comp, e.E_prev, ok);
            if (!ok) return false;
            s = s + (double)e.w;
            seq_s[seq_base[t] + k] = s;
        }
    // pass 5: apply
    for (size_t t = 0; t < nt; t++) {
        size_t a = t * TILE, b = std::min(n, a + (size_t)TILE);
        int k = 0;
        for (size_t i = a; i < b; i++) {
            if (kind[i]) { out[i] = seq_s[seq_base[t] + k]; k++; continue; }
            int rank = seq_base[t] + k;
            double start = rank ? seq_s[rank - 1] : 0.0;
            Par comp = k ? V[i] : par_compose(carry[t], V[i]);
            out[i] = par_apply(start, comp, Eof[i], ok);
            if (!ok) return false;
        }
    }
    return true;
}

int main(int argc, char** argv) {
    int rounds = argc > 1 ? atoi(argv[1]) : 40;
    std::mt19937_64 rng(12345);
    Stats st;
    long mismatches = 0;
    auto U = [&]() { return std::generate_canonical<double, 53>(rng); };
    for (int round = 0; round < rounds; round++) {
        for (int kindw = 0; kindw < 9; kindw++) {
            size_t n;
            switch (round % 5) { case 0: n = 1 + rng() % 50; break; case 1: n = 2048 + rng() % 3; break; case 2: n = 1 + rng() % 10000; break;
                                 case 3: n = 100000 + rng() % 1000; break; default: n = 300000 + rng() % 100000; }
            std::vector<float> w(n);
            for (size_t i = 0; i < n; i++) {
                double v;
                switch (kindw) {
                    case 0: v = 40.0 * U(); break;                                           // sensor-model-like weights
                    case 1: v = std::ldexp(U(), -(int)(rng() % 60) + 5); break;               // huge dynamic range
                    case 2: v = std::ldexp((double)(rng() % 64), -(int)(rng() % 40)); break;  // many ties and exact adds
                    case 3: v = (rng() % 4 == 0) ? 12.0 * U() : 0.0; break;                   // mostly zeros (invalid particles)
                    case 4: v = (i < n / 3) ? 0.0 : U(); break;                               // leading zeros
                    case 5: v = (i == 7 % n) ? 1e6 : std::ldexp(U(), -30); break;             // one giant then dust
                    case 6: v = U() / (double)n; break;                                       // normalised weights ~1/n
                    case 7: v = (rng() % 3 == 0) ? std::ldexp(1.0, -(int)(rng() % 50)) : 3.0 * U(); break;   // powers of two mixed in
                    default: v = std::ldexp(1.0 + (double)(rng() % 3), -24 - (int)(rng() % 6)); break;      // half-ulp sized terms
                }
                w[i] = (float)v;
            }
            std::vector<double> ref(n);
            double s = 0;
            for (size_t i = 0; i < n; i++) { s = s + (double)w[i]; ref[i] = s; }
            for (int order = 0; order < 3; order++) {
                std::vector<double> out;
                st.cases++;
                if (!exact_scan(w, out, order, st)) { st.fallback++; continue; }
                for (size_t i = 0; i < n; i++)
                    if (f64_bits(out[i]) != f64_bits(ref[i])) {
                        if (mismatches < 10) printf("MISMATCH kind %d n %zu i %zu got %.17g want %.17g\n", kindw, n, i, out[i], ref[i]);
                        mismatches++;
                        break;
                    }
            }
        }
    }
    printf("cases %ld fallbacks %ld seq-elements %ld uniform-tiles %ld mismatching-cases %ld\n", st.cases, st.fallback, st.seq, st.uniform_tiles, mismatches);
    return mismatches ? 1 : 0;
}
