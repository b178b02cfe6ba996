// chain.cuh -- per-pixel float64 evaluation of a lowered transformer chain (device side).
//
// Replaces the whole-image NumPy passes of the reference's MultiTransformer.transform
// (/root/reference/src/vr180_convert/transformer.py:93-98) for the recognised transformer classes.  Instead of
// materialising ~10 float64 (H,W) temporaries, each thread carries one pixel through the op list in registers.
//
// The reference converts between three coordinate forms at every step (x,y) -> (theta, roll) -> (x,y) or
// (x,y) -> unit vector -> (x,y) using atan2/sin/cos pairs.  Those round trips are algebraic identities
// (sin(atan2(a,b)) == a/hypot(a,b)), so the evaluator keeps a lazily-converted state instead:
//
//      XY     plain coordinates (x, y)
//      POLAR  signed radius r and unit direction (ux, uy):   (x, y) == r * (ux, uy)
//      VEC3   unit vector (vx, vy, vz), z forward            (transformer.py:483-530)
//
// and converts only when the next op needs another form.  This leaves one acos + one sqrt + two divisions per
// 3-D step instead of ~10 transcendental calls, and differs from the reference only by a few float64 roundings
// (~1e-13 px), i.e. the float32 map is the same float32 except for a vanishing number of round-to-nearest ties
// (SURVEY.md §7 hard part 1, Appendix A).  The one exception is a projection's singular point falling exactly on an
// output pixel (the stereographic antipode, 90 degrees off axis of the rectilinear mapping): one coordinate is ~1e17 px
// and the other the product of a ~1e-17 direction cosine and that radius, so the two algebraically equal forms differ
// by tens of pixels there; under the default constant border the pixel is the border colour either way
// (tests/test_gpu_fuzz.py).  Operations the reference performs as separate NumPy ufuncs
// (mul then add, Horner steps) use explicit __dmul_rn/__dadd_rn so no FMA contraction changes a rounding.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/vr180_b200.h"

namespace vr180 {

constexpr double kHalfPi = 1.5707963267948966;    // np.pi / 2
constexpr double kSqrt2 = 1.4142135623730951;     // np.sqrt(2)
constexpr double kTwoOverPi = 0.63661977236758134;  // 1 / (np.pi / 2)

struct ChainState {
    int mode;  // 0 XY, 1 POLAR, 2 VEC3
    double x, y;        // XY
    double r, ux, uy;   // POLAR
    double vx, vy, vz;  // VEC3
};

enum { MODE_XY = 0, MODE_POLAR = 1, MODE_VEC3 = 2 };

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// VEC3 -> POLAR: equidistant_from_3d (transformer.py:526-529): theta = arccos(vz), phi = atan2(vx, vy),
// (x, y) = theta * (sin phi, cos phi) == theta * (vx, vy) / hypot(vx, vy).
__device__ __forceinline__ void vec3_to_polar(ChainState& s) {
    const double h2 = add_rn(mul_rn(s.vx, s.vx), mul_rn(s.vy, s.vy));
    s.r = acos(s.vz);  // NaN when rounding pushed |vz| above 1, as np.arccos does
    if (h2 > 0.0) {
        const double inv = rsqrt(h2);  // 1 / hypot(vx, vy): one MUFU + Newton steps instead of sqrt + division
        s.ux = s.vx * inv;
        s.uy = s.vy * inv;
    } else {  // atan2(0, 0) = 0 -> (sin, cos) = (0, 1)
        s.ux = 0.0;
        s.uy = 1.0;
    }
    s.mode = MODE_POLAR;
}

__device__ __forceinline__ void polar_to_xy(ChainState& s) {
    s.x = mul_rn(s.r, s.ux);
    s.y = mul_rn(s.r, s.uy);
    s.mode = MODE_XY;
}

// (x,y) -> (theta >= 0, unit direction): PolarRollTransformer.transform, transformer.py:271-272
// theta = sqrt(x**2 + y**2), roll = atan2(y, x); (cos roll, sin roll) == (x, y) / theta, roll(0,0) = 0.
__device__ __forceinline__ void xy_to_polar(ChainState& s) {
    const double t = sqrt(add_rn(mul_rn(s.x, s.x), mul_rn(s.y, s.y)));
    if (t > 0.0) {
        const double inv = 1.0 / t;
        s.ux = s.x * inv;
        s.uy = s.y * inv;
    } else if (t == 0.0) {
        s.ux = 1.0;
        s.uy = 0.0;
    } else {  // NaN
        s.ux = t;
        s.uy = t;
    }
    s.r = t;
    s.mode = MODE_POLAR;
}

// Bring the state to POLAR with r >= 0, as every PolarRollTransformer sees it.
__device__ __forceinline__ void to_polar_nonneg(ChainState& s) {
    if (s.mode == MODE_VEC3) vec3_to_polar(s);
    else if (s.mode == MODE_XY) xy_to_polar(s);
    if (s.r < 0.0) {
        s.r = -s.r;
        s.ux = -s.ux;
        s.uy = -s.uy;
    } else if (s.r == 0.0) {  // the reference re-derives roll = atan2(0, 0) = 0 from (0, 0)
        s.ux = 1.0;
        s.uy = 0.0;
    }
}

__device__ __forceinline__ void to_xy(ChainState& s) {
    if (s.mode == MODE_VEC3) vec3_to_polar(s);
    if (s.mode == MODE_POLAR) polar_to_xy(s);
}

// -> VEC3: equidistant_to_3d (transformer.py:502-507): phi = atan2(x, y), theta = hypot(x, y),
// v = (sin theta sin phi, sin theta cos phi, cos theta); (sin phi, cos phi) == (x, y)/theta, phi(0,0) = 0.
__device__ __forceinline__ void to_vec3(ChainState& s) {
    if (s.mode == MODE_VEC3) return;
    if (s.mode == MODE_XY) xy_to_polar(s);
    double t = s.r, ux = s.ux, uy = s.uy;
    if (t < 0.0) { t = -t; ux = -ux; uy = -uy; }
    if (t == 0.0) { ux = 0.0; uy = 1.0; }
    double st, ct;
    sincos(t, &st, &ct);
    s.vx = mul_rn(st, ux);
    s.vy = mul_rn(st, uy);
    s.vz = ct;
    s.mode = MODE_VEC3;
}

__device__ __forceinline__ double fisheye_r_to_theta(int mapping, double r) {  // transformer.py:363-372
    switch (mapping) {
        case VR180_MAP_RECTILINEAR: return atan(r);
        case VR180_MAP_STEREOGRAPHIC: return mul_rn(2.0, atan(r));
        case VR180_MAP_EQUIDISTANT: return mul_rn(r, kHalfPi);
        case VR180_MAP_EQUISOLID: return mul_rn(2.0, asin(__ddiv_rn(r, kSqrt2)));
        default: return asin(r);  // orthographic; NaN for r > 1 like np.arcsin
    }
}

__device__ __forceinline__ double fisheye_theta_to_r(int mapping, double t) {  // transformer.py:383-392
    switch (mapping) {
        case VR180_MAP_RECTILINEAR: return tan(t);
        case VR180_MAP_STEREOGRAPHIC: return mul_rn(2.0, tan(mul_rn(t, 0.5)));
        case VR180_MAP_EQUIDISTANT: return mul_rn(t, kTwoOverPi);  // theta / (pi/2) to 1 ulp, without the division
        case VR180_MAP_EQUISOLID: return mul_rn(kSqrt2, sin(mul_rn(t, 0.5)));
        default: return sin(t);
    }
}

// ---- one function per op, shared by the interpreter below and the straight-line evaluator of tiled.cu -------
__device__ __forceinline__ void op_normalize(const double* p, ChainState& s) {  // (x - cx) / scale * 2
    to_xy(s);
    s.x = mul_rn(__ddiv_rn(add_rn(s.x, -p[0]), p[2]), 2.0);
    s.y = mul_rn(__ddiv_rn(add_rn(s.y, -p[1]), p[2]), 2.0);
}
__device__ __forceinline__ void op_denormalize(const double* p, ChainState& s) {  // x * sx + cx
    to_xy(s);
    s.x = add_rn(mul_rn(s.x, p[0]), p[2]);
    s.y = add_rn(mul_rn(s.y, p[1]), p[3]);
}
__device__ __forceinline__ void op_equirect_enc(int lat_is_y, ChainState& s) {  // transformer.py:545-566
    to_xy(s);
    const double lat = mul_rn(lat_is_y ? s.y : s.x, kHalfPi);
    const double lon = mul_rn(lat_is_y ? s.x : s.y, kHalfPi);
    double sl, cl, sn, cn;
    sincos(lat, &sl, &cl);
    sincos(lon, &sn, &cn);
    if (lat_is_y) { s.vx = mul_rn(cl, sn); s.vy = sl; }
    else          { s.vx = sl; s.vy = mul_rn(cl, sn); }
    s.vz = mul_rn(cl, cn);
    s.mode = MODE_VEC3;
}
__device__ __forceinline__ void op_rot3(const double* R, ChainState& s) {  // v' = R v (quaternion.rotate_vectors)
    to_vec3(s);
    const double a = s.vx, b = s.vy, c3 = s.vz;
    s.vx = add_rn(add_rn(mul_rn(R[0], a), mul_rn(R[1], b)), mul_rn(R[2], c3));
    s.vy = add_rn(add_rn(mul_rn(R[3], a), mul_rn(R[4], b)), mul_rn(R[5], c3));
    s.vz = add_rn(add_rn(mul_rn(R[6], a), mul_rn(R[7], b)), mul_rn(R[8], c3));
}
// np.polyval(np.flip(coefs_reverse), theta): y = y*x + c, highest power first
__device__ __forceinline__ void op_poly(const double* coefs, int n, ChainState& s) {
    to_polar_nonneg(s);
    // One uniform jump into a fully unrolled Horner chain: every coefficient is a static constant-bank operand
    // and a step is exactly DMUL + DADD (a loop over the runtime n costs ~6 uniform address instructions per step
    // on top of the two FP64 ones).  n <= VR180_MAX_OP_PARAMS is validated on the host.
    static_assert(VR180_MAX_OP_PARAMS == 12, "extend the chain below");
    const double r = s.r;
    double acc = 0.0;
#define VR180_HORNER(i) acc = add_rn(mul_rn(acc, r), coefs[i]);
    switch (n) {
        default: VR180_HORNER(11)
        case 11: VR180_HORNER(10)
        case 10: VR180_HORNER(9)
        case 9: VR180_HORNER(8)
        case 8: VR180_HORNER(7)
        case 7: VR180_HORNER(6)
        case 6: VR180_HORNER(5)
        case 5: VR180_HORNER(4)
        case 4: VR180_HORNER(3)
        case 3: VR180_HORNER(2)
        case 2: VR180_HORNER(1)
        case 1: VR180_HORNER(0)
        case 0: break;
    }
#undef VR180_HORNER
    s.r = acc;
}
__device__ __forceinline__ void op_fisheye_dec(int mapping, ChainState& s) {
    to_polar_nonneg(s);
    s.r = fisheye_theta_to_r(mapping, s.r);
}

// Apply ops [first, last) of the chain to the state.  Control flow is uniform across the grid (the op list is
// a kernel constant), so the switch costs no divergence.
__device__ __forceinline__ void run_ops(const vr180_chain_t& c, int first, int last, ChainState& s) {
    for (int k = first; k < last; ++k) {
        const vr180_op_t& op = c.ops[k];
        switch (op.code) {
            case VR180_OP_NORMALIZE: op_normalize(op.p, s); break;
            case VR180_OP_DENORMALIZE: op_denormalize(op.p, s); break;
            case VR180_OP_DENORMALIZE_INV: {  // (x - cx) / sx
                to_xy(s);
                s.x = __ddiv_rn(add_rn(s.x, -op.p[2]), op.p[0]);
                s.y = __ddiv_rn(add_rn(s.y, -op.p[3]), op.p[1]);
                break;
            }
            case VR180_OP_ZOOM: {
                if (s.mode == MODE_VEC3) vec3_to_polar(s);
                if (s.mode == MODE_POLAR) s.r = __ddiv_rn(s.r, op.p[0]);
                else { s.x = __ddiv_rn(s.x, op.p[0]); s.y = __ddiv_rn(s.y, op.p[0]); }
                break;
            }
            case VR180_OP_ZOOM_INV: {
                if (s.mode == MODE_VEC3) vec3_to_polar(s);
                if (s.mode == MODE_POLAR) s.r = mul_rn(s.r, op.p[0]);
                else { s.x = mul_rn(s.x, op.p[0]); s.y = mul_rn(s.y, op.p[0]); }
                break;
            }
            case VR180_OP_EQUIRECT_ENC: op_equirect_enc(op.iparam, s); break;
            case VR180_OP_EQUIRECT_DEC: {  // transformer.py:573-583
                to_vec3(s);
                double lat, lon;
                if (op.iparam) { lat = asin(s.vy); lon = atan2(s.vx, s.vz); s.x = __ddiv_rn(lon, kHalfPi); s.y = __ddiv_rn(lat, kHalfPi); }
                else           { lat = asin(s.vx); lon = atan2(s.vy, s.vz); s.x = __ddiv_rn(lat, kHalfPi); s.y = __ddiv_rn(lon, kHalfPi); }
                s.mode = MODE_XY;
                break;
            }
            case VR180_OP_FISHEYE_ENC: to_polar_nonneg(s); s.r = fisheye_r_to_theta(op.iparam, s.r); break;
            case VR180_OP_FISHEYE_DEC: op_fisheye_dec(op.iparam, s); break;
            case VR180_OP_RECTILINEAR_DEC: to_polar_nonneg(s); s.r = mul_rn(tan(s.r), op.p[0]); break;
            case VR180_OP_RECTILINEAR_DEC_INV: to_polar_nonneg(s); s.r = atan(__ddiv_rn(s.r, op.p[0])); break;
            case VR180_OP_POLY: op_poly(op.p, op.iparam, s); break;
            case VR180_OP_ROT3: op_rot3(op.p, s); break;
            default: break;
        }
    }
}

// Full chain for output pixel (col i, row j): meshgrid(arange(W), arange(H)) -> chain -> (xs, ys) in float64.
__device__ __forceinline__ void eval_chain(const vr180_chain_t& c, int i, int j, double& xs, double& ys) {
    ChainState s;
    s.mode = MODE_XY;
    s.x = (double)i;
    s.y = (double)j;
    s.r = s.ux = s.uy = s.vx = s.vy = s.vz = 0.0;
    run_ops(c, 0, c.n_ops, s);
    to_xy(s);
    xs = s.x;
    ys = s.y;
}

// Chain WITHOUT its final DENORMALIZE op: the frame-independent normalised source position.
__device__ __forceinline__ void eval_chain_normalised(const vr180_chain_t& c, int i, int j, double& nx, double& ny) {
    ChainState s;
    s.mode = MODE_XY;
    s.x = (double)i;
    s.y = (double)j;
    s.r = s.ux = s.uy = s.vx = s.vy = s.vz = 0.0;
    run_ops(c, 0, c.n_ops - 1, s);
    to_xy(s);
    nx = s.x;
    ny = s.y;
}

}  // namespace vr180
