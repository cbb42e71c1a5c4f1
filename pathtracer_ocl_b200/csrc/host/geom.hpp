// geom.hpp -- host-side tuple / 4x4 matrix math used while building scenes.
//
// Restates the conventions of the reference's internal/app/geom package so the scene buffers
// this frontend produces are numerically the ones the Go frontend would produce:
//   - row-major 4x4, translation in elements [3],[7],[11]      (geom/translation.go:5-12)
//   - products and matrix*tuple sum a+b+c+d left to right       (geom/matrix.go:50-72, 215-222)
//   - inverse through cofactors / determinant                   (geom/matrix.go:200-213)
//   - magnitude over x,y,z only, normalise divides all four     (geom/tuple.go:162-185)
// Go on amd64 never fuses multiply-add, so this file must be compiled with -ffp-contract=off.
#pragma once
#include <array>
#include <cmath>

namespace pt {

using Tuple4 = std::array<double, 4>;
using Mat4 = std::array<double, 16>;

inline Tuple4 point(double x, double y, double z) { return {x, y, z, 1.0}; }
inline Tuple4 vector(double x, double y, double z) { return {x, y, z, 0.0}; }
inline Tuple4 color(double r, double g, double b) { return {r, g, b, 1.0}; }   // geom.NewColor
inline Tuple4 rgb0(double r, double g, double b) { return {r, g, b, 0.0}; }    // Tuple4{r,g,b} literal

inline Tuple4 add(const Tuple4& a, const Tuple4& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]}; }
inline Tuple4 sub(const Tuple4& a, const Tuple4& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]}; }
inline double magnitude(const Tuple4& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline Tuple4 normalize(const Tuple4& a) {
    double m = magnitude(a);
    return {a[0] / m, a[1] / m, a[2] / m, a[3] / m};
}
inline Tuple4 cross(const Tuple4& a, const Tuple4& b) {
    return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0], 0.0};
}
// geom.Eq / TupleEquals: |a-b| < 0.01 on all four components (geom/types.go:5-12, tuple.go:258-263)
inline bool approx(double a, double b) { return std::fabs(a - b) < 0.01; }
inline bool tuple_equals(const Tuple4& a, const Tuple4& b) {
    return approx(a[0], b[0]) && approx(a[1], b[1]) && approx(a[2], b[2]) && approx(a[3], b[3]);
}

inline Mat4 identity() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; }

inline Mat4 multiply(const Mat4& a, const Mat4& b) {
    Mat4 m{};
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            double p0 = a[r * 4 + 0] * b[0 + c];
            double p1 = a[r * 4 + 1] * b[4 + c];
            double p2 = a[r * 4 + 2] * b[8 + c];
            double p3 = a[r * 4 + 3] * b[12 + c];
            m[r * 4 + c] = p0 + p1 + p2 + p3;
        }
    return m;
}

inline Tuple4 multiply(const Mat4& m, const Tuple4& t) {
    Tuple4 o{};
    for (int r = 0; r < 4; ++r) {
        double p0 = m[r * 4 + 0] * t[0];
        double p1 = m[r * 4 + 1] * t[1];
        double p2 = m[r * 4 + 2] * t[2];
        double p3 = m[r * 4 + 3] * t[3];
        o[r] = p0 + p1 + p2 + p3;
    }
    return o;
}

inline Mat4 transpose(const Mat4& m) {
    Mat4 o{};
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) o[r * 4 + c] = m[c * 4 + r];
    return o;
}

namespace detail {
inline double det2(const double* m) { return m[0] * m[3] - m[1] * m[2]; }
inline void sub3(const double* m3, int dr, int dc, double* out2) {
    int k = 0;
    for (int r = 0; r < 3; ++r) {
        if (r == dr) continue;
        for (int c = 0; c < 3; ++c) {
            if (c == dc) continue;
            out2[k++] = m3[r * 3 + c];
        }
    }
}
inline double cofactor3(const double* m3, int r, int c) {
    double s[4];
    sub3(m3, r, c, s);
    double minor = det2(s);
    return ((r + c) % 2 != 0) ? -minor : minor;
}
inline double det3(const double* m3) {
    double d = 0.0;
    for (int c = 0; c < 3; ++c) d = d + m3[c] * cofactor3(m3, 0, c);
    return d;
}
inline void sub4(const Mat4& m, int dr, int dc, double* out3) {
    int k = 0;
    for (int r = 0; r < 4; ++r) {
        if (r == dr) continue;
        for (int c = 0; c < 4; ++c) {
            if (c == dc) continue;
            out3[k++] = m[r * 4 + c];
        }
    }
}
}  // namespace detail

inline double cofactor4(const Mat4& m, int r, int c) {
    double s[9];
    detail::sub4(m, r, c, s);
    double minor = detail::det3(s);
    return ((r + c) % 2 != 0) ? -minor : minor;
}
inline double determinant(const Mat4& m) {
    double d = 0.0;
    for (int c = 0; c < 4; ++c) d = d + m[c] * cofactor4(m, 0, c);
    return d;
}
inline Mat4 inverse(const Mat4& m) {
    Mat4 o{};
    double d = determinant(m);
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) o[c * 4 + r] = cofactor4(m, r, c) / d;
    return o;
}

inline Mat4 translate(double x, double y, double z) {
    Mat4 m = identity();
    m[3] = x; m[7] = y; m[11] = z;
    return m;
}
inline Mat4 scale(double x, double y, double z) {
    Mat4 m = identity();
    m[0] = x; m[5] = y; m[10] = z;
    return m;
}
inline Mat4 rotate_x(double r) {
    Mat4 m = identity();
    m[5] = std::cos(r); m[6] = -std::sin(r); m[9] = std::sin(r); m[10] = std::cos(r);
    return m;
}
inline Mat4 rotate_y(double r) {
    Mat4 m = identity();
    m[0] = std::cos(r); m[2] = std::sin(r); m[8] = -std::sin(r); m[10] = std::cos(r);
    return m;
}
inline Mat4 rotate_z(double r) {
    Mat4 m = identity();
    m[0] = std::cos(r); m[1] = -std::sin(r); m[4] = std::sin(r); m[5] = std::cos(r);
    return m;
}

}  // namespace pt
