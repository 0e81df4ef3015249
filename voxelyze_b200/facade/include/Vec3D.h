// Vec3D.h -- value type used in the signatures of the drop-in C++ API (voxelyze_b200 facade).
//
// Same public surface as the reference's include/Vec3D.h (x/y/z members, arithmetic operators,
// Length/Dot/Cross/Normalized ...) so that caller code compiles unchanged; written from scratch.
// Note the reference's division multiplies by the reciprocal (include/Vec3D.h:65) and the facade
// keeps that rounding behaviour where callers could observe it.
#ifndef VXB200_VEC3D_H
#define VXB200_VEC3D_H

#include <math.h>      // like the reference (include/Vec3D.h:17-18): callers rely on the global abs(double) / sqrt overloads it brings
#include <float.h>
#include <cmath>

#define vec3_X 0
#define vec3_Y 1
#define vec3_Z 2

template <typename T = double>
class Vec3D {
public:
    T x, y, z;

    Vec3D() : x(0), y(0), z(0) {}
    Vec3D(T ax, T ay, T az) : x(ax), y(ay), z(az) {}
    Vec3D(const Vec3D& o) : x(o.x), y(o.y), z(o.z) {}
    template <typename U> Vec3D(const Vec3D<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}

    Vec3D& operator=(const Vec3D& o) { x = o.x; y = o.y; z = o.z; return *this; }
    template <typename U> Vec3D& operator=(const Vec3D<U>& o) { x = (T)o.x; y = (T)o.y; z = (T)o.z; return *this; }
    template <typename U> operator Vec3D<U>() const { return Vec3D<U>((U)x, (U)y, (U)z); }

    bool IsValid() const { return std::isfinite((double)x) && std::isfinite((double)y) && std::isfinite((double)z); }

    // arithmetic
    Vec3D operator+(const Vec3D& o) const { return Vec3D(x + o.x, y + o.y, z + o.z); }
    Vec3D operator-(const Vec3D& o) const { return Vec3D(x - o.x, y - o.y, z - o.z); }
    Vec3D operator-() const { return Vec3D(-x, -y, -z); }
    Vec3D operator*(const T& f) const { return Vec3D(f * x, f * y, f * z); }
    Vec3D operator/(const T& f) const { T inv = (T)1.0 / f; return Vec3D(inv * x, inv * y, inv * z); }
    friend Vec3D operator*(const T f, const Vec3D& v) { return v * f; }
    template <typename U> Vec3D operator+(const Vec3D<U>& o) const { return Vec3D(x + o.x, y + o.y, z + o.z); }
    template <typename U> Vec3D operator-(const Vec3D<U>& o) const { return Vec3D(x - o.x, y - o.y, z - o.z); }
    Vec3D& operator+=(const Vec3D& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vec3D& operator-=(const Vec3D& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    template <typename U> Vec3D& operator+=(const Vec3D<U>& o) { x += o.x; y += o.y; z += o.z; return *this; }
    template <typename U> Vec3D& operator-=(const Vec3D<U>& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    Vec3D& operator*=(const T f) { x *= f; y *= f; z *= f; return *this; }
    Vec3D& operator/=(const T f) { T inv = (T)1.0 / f; x *= inv; y *= inv; z *= inv; return *this; }
    bool operator==(const Vec3D& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const Vec3D& o) const { return !(*this == o); }

    const T& operator[](int i) const { int k = i % 3; return k == 0 ? x : (k == 1 ? y : z); }
    T& operator[](int i) { int k = i % 3; return k == 0 ? x : (k == 1 ? y : z); }

    T getX() const { return x; }
    T getY() const { return y; }
    T getZ() const { return z; }
    void setX(T v) { x = v; }
    void setY(T v) { y = v; }
    void setZ(T v) { z = v; }

    // in-place operations
    T Normalize()
    {
        T l = std::sqrt(x * x + y * y + z * z);
        if (l > 0) { x /= l; y /= l; z /= l; }
        return l;
    }
    void NormalizeFast()
    {
        T l = std::sqrt(x * x + y * y + z * z);
        if (l > 0) { T li = (T)1.0 / l; x *= li; y *= li; z *= li; }
    }
    void RotZ(T a) { T c = std::cos(a), s = std::sin(a); T nx = x * c - y * s, ny = x * s + y * c; x = nx; y = ny; }
    void RotY(T a) { T c = std::cos(a), s = std::sin(a); T nx = x * c + z * s, nz = -x * s + z * c; x = nx; z = nz; }
    void RotX(T a) { T c = std::cos(a), s = std::sin(a); T ny = y * c + z * s, nz = -y * s + z * c; y = ny; z = nz; }

    // pure functions
    Vec3D Cross(const Vec3D& v) const { return Vec3D(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x); }
    T Dot(const Vec3D& v) const { return x * v.x + y * v.y + z * v.z; }
    Vec3D Abs() const { return Vec3D(x >= 0 ? x : -x, y >= 0 ? y : -y, z >= 0 ? z : -z); }
    Vec3D Normalized() const { T l = std::sqrt(x * x + y * y + z * z); return l > 0 ? (*this) / l : (*this); }
    bool IsNear(const Vec3D& s, T thresh) const { return Dist2(s) < thresh * thresh; }
    T Length() const { return std::sqrt(x * x + y * y + z * z); }
    T Length2() const { return x * x + y * y + z * z; }
    Vec3D Min(const Vec3D& s) const { return Vec3D(x < s.x ? x : s.x, y < s.y ? y : s.y, z < s.z ? z : s.z); }
    Vec3D Max(const Vec3D& s) const { return Vec3D(x > s.x ? x : s.x, y > s.y ? y : s.y, z > s.z ? z : s.z); }
    T Min() const { T m = x < y ? x : y; return z < m ? z : m; }
    T Max() const { T m = x > y ? x : y; return z > m ? z : m; }
    Vec3D Scale(const Vec3D& v) const { return Vec3D(x * v.x, y * v.y, z * v.z); }
    Vec3D ScaleInv(const Vec3D& v) const { return Vec3D(x / v.x, y / v.y, z / v.z); }
    T Dist(const Vec3D& v) const { return std::sqrt(Dist2(v)); }
    T Dist2(const Vec3D& v) const { T dx = v.x - x, dy = v.y - y, dz = v.z - z; return dx * dx + dy * dy + dz * dz; }
};

#endif // VXB200_VEC3D_H
