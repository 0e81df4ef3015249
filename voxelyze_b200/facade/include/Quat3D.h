// Quat3D.h -- quaternion value type of the drop-in C++ API (voxelyze_b200 facade).
//
// Public surface of the reference's include/Quat3D.h (w/x/y/z members, product, conjugate,
// rotation-vector conversions with the reference's thresholds, vector rotation); written from
// scratch.  Thresholds: include/Quat3D.h:19-29 of the reference.
#ifndef VXB200_QUAT3D_H
#define VXB200_QUAT3D_H

#include <math.h>      // like the reference (include/Vec3D.h:17-18): callers rely on the global abs(double) / sqrt overloads it brings
#include <float.h>
#include <cmath>
#include "Vec3D.h"

template <typename T = double>
class Quat3D {
public:
    T w, x, y, z;

    Quat3D() : w(1), x(0), y(0), z(0) {}
    Quat3D(T aw, T ax, T ay, T az) : w(aw), x(ax), y(ay), z(az) {}
    Quat3D(const Quat3D& q) : w(q.w), x(q.x), y(q.y), z(q.z) {}
    Quat3D(const Vec3D<T>& rotationVector) { FromRotationVector(rotationVector); }   // rotation vector, like the reference
    Quat3D(T angle, const Vec3D<T>& axis)
    {
        T h = angle * (T)0.5, s = std::sin(h);
        w = std::cos(h); x = axis.x * s; y = axis.y * s; z = axis.z * s;
    }
    template <typename U> Quat3D(const Quat3D<U>& q) : w((T)q.w), x((T)q.x), y((T)q.y), z((T)q.z) {}
    template <typename U> operator Quat3D<U>() const { return Quat3D<U>((U)w, (U)x, (U)y, (U)z); }

    Quat3D& operator=(const Quat3D& q) { w = q.w; x = q.x; y = q.y; z = q.z; return *this; }
    Quat3D operator+(const Quat3D& q) const { return Quat3D(w + q.w, x + q.x, y + q.y, z + q.z); }
    Quat3D operator-(const Quat3D& q) const { return Quat3D(w - q.w, x - q.x, y - q.y, z - q.z); }
    Quat3D operator*(const T f) const { return Quat3D(w * f, x * f, y * f, z * f); }
    friend Quat3D operator*(const T f, const Quat3D& q) { return q * f; }
    // Hamilton product, term order of the reference (include/Quat3D.h:83)
    Quat3D operator*(const Quat3D& q) const
    {
        return Quat3D(w * q.w - x * q.x - y * q.y - z * q.z,
                      w * q.x + x * q.w + y * q.z - z * q.y,
                      w * q.y - x * q.z + y * q.w + z * q.x,
                      w * q.z + x * q.y - y * q.x + z * q.w);
    }
    bool operator==(const Quat3D& q) const { return w == q.w && x == q.x && y == q.y && z == q.z; }
    bool operator!=(const Quat3D& q) const { return !(*this == q); }
    Quat3D& operator+=(const Quat3D& q) { w += q.w; x += q.x; y += q.y; z += q.z; return *this; }
    Quat3D& operator-=(const Quat3D& q) { w -= q.w; x -= q.x; y -= q.y; z -= q.z; return *this; }

    Vec3D<T> ToVec() const { return Vec3D<T>(x, y, z); }
    T Length() const { return std::sqrt(Length2()); }
    T Length2() const { return w * w + x * x + y * y + z * z; }
    T Normalize()
    {
        T l = Length();
        if (l == 0) { w = 1; x = y = z = 0; }
        else if (l > 0) { T li = (T)1.0 / l; w *= li; x *= li; y *= li; z *= li; }
        return l;
    }
    void NormalizeFast()
    {
        T l = std::sqrt(x * x + y * y + z * z + w * w);
        if (l != 0) { T li = (T)1.0 / l; w *= li; x *= li; y *= li; z *= li; }
        if (w >= 1.0) { w = 1.0; x = y = z = 0; }
    }
    Quat3D Inverse() const { T n = Length2(); return Quat3D(w / n, -x / n, -y / n, -z / n); }
    Quat3D Conjugate() const { return Quat3D(w, -x, -y, -z); }

    T Angle() const { return (T)2.0 * std::acos(w > 1 ? (T)1 : w); }
    T AngleDegrees() const { return Angle() * (T)57.29577951308232; }
    bool IsNegligibleAngle() const { return 2.0 * std::acos(w) < 1e-7; }
    bool IsSmallAngle() const { return w > 0.9999625; }
    Vec3D<T> Axis() const
    {
        T sl = (T)1.0 - w * w;
        if (sl <= 0) return Vec3D<T>(1, 0, 0);
        return Vec3D<T>(x, y, z) / std::sqrt(sl);
    }
    Vec3D<T> AxisUnNormalized() const { return Vec3D<T>(x, y, z); }
    void AngleAxisUnNormalized(T& angle, Vec3D<T>& axis) const
    {
        if (w >= 1.0) { angle = 0; axis = Vec3D<T>(1, 0, 0); return; }
        angle = (T)2.0 * std::acos(w > 1 ? (T)1 : w);
        axis = Vec3D<T>(x, y, z);
    }
    void AngleAxis(T& angle, Vec3D<T>& axis) const { AngleAxisUnNormalized(angle, axis); axis.NormalizeFast(); }

    // quaternion -> rotation vector; sqrt approximation of acos below squareLength 2.4e-3
    Vec3D<T> ToRotationVector() const
    {
        if (w >= 1.0 || w <= -1.0) return Vec3D<T>(0, 0, 0);
        T sl = (T)1.0 - w * w;
        Vec3D<T> twice = Vec3D<T>(x, y, z) * (T)2.0;
        if (sl < 2.4e-3) return twice * std::sqrt((2 - 2 * w) / sl);
        return (twice * std::acos(w)) / std::sqrt(sl);
    }
    // rotation vector -> quaternion; Taylor branch when the 4th-order term is below 24*DBL_EPSILON
    void FromRotationVector(const Vec3D<T>& v)
    {
        Vec3D<T> h = v / (T)2;
        T m2 = h.Length2(), s;
        if (m2 * m2 < 5.328e-15) { w = (T)1.0 - (T)0.5 * m2; s = (T)1.0 - m2 / (T)6.0; }
        else { T m = std::sqrt(m2); w = std::cos(m); s = std::sin(m) / m; }
        x = h.x * s; y = h.y * s; z = h.z * s;
    }
    // rotation that turns `from` onto +X
    void FromAngleToPosX(const Vec3D<T>& from)
    {
        if (from.x == 0 && from.y == 0 && from.z == 0) return;
        T yox = from.y / from.x, zox = from.z / from.x;
        const T sa = (T)1.732e-2;
        if (yox < sa && yox > -sa && zox < sa && zox > -sa) {
            x = 0; y = (T)0.5 * zox; z = (T)-0.5 * yox;
            w = 1 + (T)0.5 * (-y * y - z * z);
            return;
        }
        Vec3D<T> n = from; n.NormalizeFast();
        T theta = std::acos(n.x);
        if (theta > (T)3.14159265358979 - (T)1e-7) { w = 0; x = 0; y = 1; z = 0; return; }
        T ami = (T)1.0 / std::sqrt(n.z * n.z + n.y * n.y);
        T a = (T)0.5 * theta, s = std::sin(a);
        w = std::cos(a); x = 0; y = n.z * ami * s; z = -n.y * ami * s;
    }

    Vec3D<T> RotateVec3D(const Vec3D<T>& f) const
    {
        T tw = f.x * x + f.y * y + f.z * z;
        T tx = f.x * w - f.y * z + f.z * y;
        T ty = f.x * z + f.y * w - f.z * x;
        T tz = -f.x * y + f.y * x + f.z * w;
        return Vec3D<T>(w * tx + x * tw + y * tz - z * ty, w * ty - x * tz + y * tw + z * tx, w * tz + x * ty - y * tx + z * tw);
    }
    template <typename U> Vec3D<U> RotateVec3D(const Vec3D<U>& f) const { return Vec3D<U>(RotateVec3D(Vec3D<T>(f))); }
    Vec3D<T> RotateVec3DInv(const Vec3D<T>& f) const
    {
        T tw = x * f.x + y * f.y + z * f.z;
        T tx = w * f.x - y * f.z + z * f.y;
        T ty = w * f.y + x * f.z - z * f.x;
        T tz = w * f.z - x * f.y + y * f.x;
        return Vec3D<T>(tw * x + tx * w + ty * z - tz * y, tw * y - tx * z + ty * w + tz * x, tw * z + tx * y - ty * x + tz * w);
    }
    template <typename U> Vec3D<U> RotateVec3DInv(const Vec3D<U>& f) const { return Vec3D<U>(RotateVec3DInv(Vec3D<T>(f))); }
};

#endif // VXB200_QUAT3D_H
