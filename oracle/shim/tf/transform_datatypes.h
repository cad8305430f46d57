// Stand-in for ROS tf (ros/geometry, not vendored by the reference, version unpinned).
// Restates tf::Quaternion::setRPY, quaternionMsgToTF, Matrix3x3::setRotation and getEulerYPR
// in the operation order of the library. This boundary stays "parity unpinned" (SURVEY §8c).
#pragma once
#include <math.h>
#include <geometry_msgs/Pose.h>
namespace tf {
class Quaternion {
    double m[4];
public:
    Quaternion() { m[0] = m[1] = m[2] = 0; m[3] = 1; }
    Quaternion(double x, double y, double z, double w) { m[0] = x; m[1] = y; m[2] = z; m[3] = w; }
    void setRPY(double roll, double pitch, double yaw) {
        double halfYaw = yaw * 0.5, halfPitch = pitch * 0.5, halfRoll = roll * 0.5;
        double cosYaw = ::cos(halfYaw), sinYaw = ::sin(halfYaw);
        double cosPitch = ::cos(halfPitch), sinPitch = ::sin(halfPitch);
        double cosRoll = ::cos(halfRoll), sinRoll = ::sin(halfRoll);
        m[0] = sinRoll * cosPitch * cosYaw - cosRoll * sinPitch * sinYaw;
        m[1] = cosRoll * sinPitch * cosYaw + sinRoll * cosPitch * sinYaw;
        m[2] = cosRoll * cosPitch * sinYaw - sinRoll * sinPitch * cosYaw;
        m[3] = cosRoll * cosPitch * cosYaw + sinRoll * sinPitch * sinYaw;
    }
    double x() const { return m[0]; } double y() const { return m[1]; } double z() const { return m[2]; } double w() const { return m[3]; }
    double length2() const { return m[0] * m[0] + m[1] * m[1] + m[2] * m[2] + m[3] * m[3]; }
    void normalize() { double l = ::sqrt(length2()); for (double& v : m) v /= l; }
};
inline geometry_msgs::Quaternion createQuaternionMsgFromYaw(double yaw) {
    Quaternion q; q.setRPY(0.0, 0.0, yaw);
    geometry_msgs::Quaternion o; o.x = q.x(); o.y = q.y(); o.z = q.z(); o.w = q.w(); return o;
}
inline double getYaw(const geometry_msgs::Quaternion& mq) {
    Quaternion q(mq.x, mq.y, mq.z, mq.w);
    if (::fabs(q.length2() - 1) > 0.1) q.normalize();
    double d = q.length2();
    double s = 2.0 / d;
    double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    double wy = q.w() * ys, wz = q.w() * zs;
    double xy = q.x() * ys, xz = q.x() * zs;
    double yy = q.y() * ys, zz = q.z() * zs;
    double m00 = 1.0 - (yy + zz), m10 = xy + wz, m20 = xz - wy;
    if (::fabs(m20) >= 1) return 0.0;
    double pitch = -::asin(m20);
    return ::atan2(m10 / ::cos(pitch), m00 / ::cos(pitch));
}
}  // namespace tf
