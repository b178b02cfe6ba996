// chain_fast.cuh -- the standard chain (Normalize, EquirectangularEncoder, [Euclidean3DRotator], [PolynomialScaler],
// FisheyeDecoder("equidistant"), Denormalize) folded for the tiled kernel's prologue: ~85 instructions per pixel instead
// of ~230 for the op-by-op evaluation of chain.cuh.
//
// Replaces, for that chain shape, the NumPy passes of /root/reference/src/vr180_convert/transformer.py:144-170 (Normalize),
// :534-566 (EquirectangularEncoder), :676 (rotate_vectors), :511-530 (equidistant_from_3d: arccos, arctan2), :441-452
// (PolynomialScaler), :363-392 (FisheyeDecoder) and :189-204 (Denormalize).  Three algebraic foldings, each
// within a few float64 roundings of the reference's order of operations (~1e-12 px; the float32 map the reference
// rounds to has a spacing of 1.2e-4 .. 2.4e-4 px at 8K, and the 16.7 M coordinates of a 4096^2 map come out identical,
// tests/test_gpu_fast_chain.py::test_fast_chain_equals_the_op_by_op_chain; 1.14 G pixels of random cases: scripts/fast_chain_sweep.py):
//   * rotation: v = (c_row s_col, s_row, c_row c_col), so R v = c_row (R0 s_col + R2 c_col) + R1 s_row -- the bracket
//     depends on the column only and R1 s_row on the row only; both are tabulated per tile (shared memory, FP64) and a
//     pixel costs 3 DFMA instead of 3 DMUL + 9 DMUL + 6 DADD;
//   * theta = arccos(vz): the pixel also needs h = hypot(vx, vy) for its direction (vx, vy) / h, and (h, vz) = (sin, cos)
//     of theta.  With u = min(|vz|, h) <= 0.7072, i = round(128 u), the angle alpha = asin(u) is asin(i / 128) (table) +
//     asin(delta), delta = u cos_i - w sin_i (|delta| < 0.0056: a degree-7 series), and theta is alpha, pi/2 -+ alpha or
//     pi - alpha by octant.  No divergent branches (CUDA's acos takes two different paths around |x| = 0.57, and warps
//     of an equirectangular tile usually straddle it: 63 instructions), no MUFU, ~30 instructions.  As np.arccos, the
//     result is NaN when rounding pushed |vz| above 1.  Requires an orthonormal R (checked on the host: |R^T R - I| <
//     1e-12), else h^2 + vz^2 != 1 and the chain keeps the op-by-op evaluation;
//   * Horner steps, hypot and the final  r ux sx + cx  use FMA.
#pragma once
#include "chain.cuh"

namespace vr180 {

// {sqrt(1 - (i/128)^2), asin(i/128)}, i = 0 .. 91 (mpmath, 50 digits, rounded to float64)
static __device__ const double2 kAsinTab[92] = {
    {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.fffbfffbfff80p-1, 0x1.0000aaabdde0cp-7},
    {0x1.ffefffbffdfffp-1, 0x1.0002aabdde94cp-6},
    {0x1.ffdbfebbe9360p-1, 0x1.80090091d9024p-6},
    {0x1.ffbffbff7fec0p-1, 0x1.000aabde0b9c8p-5},
    {0x1.ff9bf63a1740bp-1, 0x1.4014d8ffaf8afp-5},
    {0x1.ff6febba4bfeap-1, 0x1.8024091fdb0a9p-5},
    {0x1.ff3bda6d9c950p-1, 0x1.c0393e65c2c93p-5},
    {0x1.feffbfdfebf1fp-1, 0x1.002abde953619p-4},
    {0x1.febb993aecf99p-1, 0x1.203ce2b380cd3p-4},
    {0x1.fe6f634576477p-1, 0x1.405390240e6fdp-4},
    {0x1.fe1b1a62bddadp-1, 0x1.606f49730ccc5p-4},
    {0x1.fdbeba917c3f5p-1, 0x1.809092913e52ep-4},
    {0x1.fd5a3f6af6b74p-1, 0x1.a0b7f03ba78acp-4},
    {0x1.fceda421efdb5p-1, 0x1.c0e5e80f7172dp-4},
    {0x1.fc78e3817e16ep-1, 0x1.e11b009e269b5p-4},
    {0x1.fbfbf7ebc755fp-1, 0x1.00abe0c129e1ep-3},
    {0x1.fb76db58a1299p-1, 0x1.10ce59ba4a8c4p-3},
    {0x1.fae987541497fp-1, 0x1.20f530308cc20p-3},
    {0x1.fa53f4fcc4b79p-1, 0x1.3120a9bed2f46p-3},
    {0x1.f9b61d0237250p-1, 0x1.41510cb011423p-3},
    {0x1.f90ff7a2fd4d2p-1, 0x1.5186a00ade974p-3},
    {0x1.f8617caabd6f6p-1, 0x1.61c1ab9d55d30p-3},
    {0x1.f7aaa3701a270p-1, 0x1.720278094cd3cp-3},
    {0x1.f6eb62d27730dp-1, 0x1.82494ed0e78fcp-3},
    {0x1.f623b1379a09bp-1, 0x1.92967a638db38p-3},
    {0x1.f553848924e81p-1, 0x1.a2ea462b4998ep-3},
    {0x1.f47ad231ea746p-1, 0x1.b344fe9a97c4dp-3},
    {0x1.f3998f1b1886cp-1, 0x1.c3a6f13aae84bp-3},
    {0x1.f2afafa9380f9p-1, 0x1.d4106cba45b08p-3},
    {0x1.f1bd27b9002c4p-1, 0x1.e481c0fce7134p-3},
    {0x1.f0c1ea9bfa45fp-1, 0x1.f4fb3f2ad079bp-3},
    {0x1.efbdeb14f4edap-1, 0x1.02be9ce0b87cdp-2},
    {0x1.eeb11b5442ff1p-1, 0x1.0b04025245cccp-2},
    {0x1.ed9b6cf3c4663p-1, 0x1.134dfa9805147p-2},
    {0x1.ec7cd0f2b5ae0p-1, 0x1.1b9cb12545e62p-2},
    {0x1.eb5537b1434dap-1, 0x1.23f0523c5dc2bp-2},
    {0x1.ea2490ebdd6b8p-1, 0x1.2c490af8bde81p-2},
    {0x1.e8eacbb648910p-1, 0x1.34a709597aab1p-2},
    {0x1.e7a7d6766784bp-1, 0x1.3d0a7c4c4bd9cp-2},
    {0x1.e65b9edeba38ep-1, 0x1.457393b90e2aap-2},
    {0x1.e50611e88d6b5p-1, 0x1.4de2808dce513p-2},
    {0x1.e3a71bcdd63dep-1, 0x1.565774cb66f02p-2},
    {0x1.e23ea802b4b1ap-1, 0x1.5ed2a392bb50fp-2},
    {0x1.e0cca12e97895p-1, 0x1.675441329986ep-2},
    {0x1.df50f124fba75p-1, 0x1.6fdc83364f719p-2},
    {0x1.ddcb80ddc085bp-1, 0x1.786ba074fef93p-2},
    {0x1.dc3c386d0ae09p-1, 0x1.8101d121bed2dp-2},
    {0x1.daa2fefaae1d8p-1, 0x1.899f4edc962d3p-2},
    {0x1.d8ffbab9145d4p-1, 0x1.924454c462cc4p-2},
    {0x1.d75250db9c792p-1, 0x1.9af11f89ba61cp-2},
    {0x1.d59aa58c6471cp-1, 0x1.a3a5ed82d9537p-2},
    {0x1.d3d89be176072p-1, 0x1.ac62fec0b2a92p-2},
    {0x1.d20c15d14a4e5p-1, 0x1.b5289525368abp-2},
    {0x1.d034f42698214p-1, 0x1.bdf6f47ae6904p-2},
    {0x1.ce5316736032ep-1, 0x1.c6ce628dd132cp-2},
    {0x1.cc665b0328622p-1, 0x1.cfaf27460fe9fp-2},
    {0x1.ca6e9ecc569b9p-1, 0x1.d8998cc3e6049p-2},
    {0x1.c86bbd609a260p-1, 0x1.e18ddf7da106bp-2},
    {0x1.c65d90dc509f4p-1, 0x1.ea8c6e5f5e67fp-2},
    {0x1.c443f1d4d22afp-1, 0x1.f3958aecddef4p-2},
    {0x1.c21eb7458e5ccp-1, 0x1.fca989658baafp-2},
    {0x1.bfedb67be13b3p-1, 0x1.02e46075785a1p-1},
    {0x1.bdb0c30185485p-1, 0x1.0779c5d4df4b8p-1},
    {0x1.bb67ae8584caap-1, 0x1.0c152382d7366p-1},
    {0x1.b91248c38986bp-1, 0x1.10b6a9e43942fp-1},
    {0x1.b6b05f6966b9bp-1, 0x1.155e8b2a00052p-1},
    {0x1.b441bdfab5580p-1, 0x1.1a0cfb6c3e9ebp-1},
    {0x1.b1c62db2564fep-1, 0x1.1ec230c714a96p-1},
    {0x1.af3d7561a9c43p-1, 0x1.237e6379cdfc7p-1},
    {0x1.aca7594d44cbdp-1, 0x1.2841ce0862975p-1},
    {0x1.aa039b06e926dp-1, 0x1.2d0cad5f90e20p-1},
    {0x1.a751f9447b724p-1, 0x1.31df40fbd31cdp-1},
    {0x1.a4922fb3ac8c2p-1, 0x1.36b9cb13786e1p-1},
    {0x1.a1c3f6ca01f29p-1, 0x1.3b9c90c43296dp-1},
    {0x1.9ee70390dec3dp-1, 0x1.4087da4473296p-1},
    {0x1.9bfb076d236ebp-1, 0x1.457bf318fe517p-1},
    {0x1.98ffafe1ece2fp-1, 0x1.4a792a4f26152p-1},
    {0x1.95f4a64decda8p-1, 0x1.4f7fd2bc2fb34p-1},
    {0x1.92d98fa2c355ep-1, 0x1.5490434275b92p-1},
    {0x1.8fae0c15ad38ap-1, 0x1.59aad71ced00fp-1},
    {0x1.8c71b6c8c49b4p-1, 0x1.5ecfee31c96e7p-1},
    {0x1.8924256bf4545p-1, 0x1.63ffed6d198f6p-1},
    {0x1.85c4e7d4a0bb1p-1, 0x1.693b3f244ee17p-1},
    {0x1.8253878ae2e09p-1, 0x1.6e825383cc40bp-1},
    {0x1.7ecf874b086dfp-1, 0x1.73d5a107bde74p-1},
    {0x1.7b386279d7bf3p-1, 0x1.7935a501afa78p-1},
    {0x1.778d8c89dc27cp-1, 0x1.7ea2e42c9027ap-1},
    {0x1.73ce704fb7b23p-1, 0x1.841deb5114bb4p-1},
    {0x1.6ffa6f4323c0dp-1, 0x1.89a74ffcc34a4p-1},
    {0x1.6c10e0a9e5d65p-1, 0x1.8f3fb14e496b4p-1},
    {0x1.681110a985d4dp-1, 0x1.94e7b8da3cf7ap-1},
};
constexpr double kPio2Hi = 0x1.921fb54442d18p+0, kPio2Lo = 0x1.1a62633145c07p-54;  // pi / 2 = hi + lo

// theta = atan2(h, vz) for (h, vz) on the unit circle, h >= 0  (== arccos(vz), see above)
__device__ __forceinline__ double theta_unit(double vz, double h) {
    const double az = fabs(vz);
    const bool swap = az > h;  // the smaller of the two is the sine of alpha <= pi / 4
    const double u = swap ? h : az, w = swap ? az : h;
    const int i = min(__double2int_rn(u * 128.0), 91);  // NaN -> 0
    const double2 t = __ldg(kAsinTab + i);
    const double d = fma(u, t.x, -(w * ((double)i * 0.0078125)));  // sin(alpha - alpha_i)
    const double d2 = d * d;
    double p = fma(d2, 15.0 / 336.0, 3.0 / 40.0);
    p = fma(d2, p, 1.0 / 6.0);
    const double al = t.y + fma(d, p * d2, d);
    const bool neg = vz < 0.0;
    const double k = swap ? (neg ? 2.0 : 0.0) : 1.0;  // theta = k pi/2 +- alpha
    double th = fma(k, kPio2Hi, swap == neg ? -al : al);
    th = fma(k, kPio2Lo, th);
    return az <= 1.0 ? th : CUDART_NAN;
}

// Per-tile tables of the folded rotation (see above): side X carries {scale, B0, B1, B2}, side Y carries {A0, A1, A2};
// X = rows when the latitude runs along y (EquirectangularEncoder's default), else columns.
//   v = scale A + B
constexpr int kStdTableDoubles = 7 * 32;  // x_side[4][32] (scale, B0, B1, B2), y_side[3][32] (A0, A1, A2)
__device__ __forceinline__ void std_tables(const double* R, bool lat_is_y, bool is_row, double sv, double cv, double* x_side,
                                           double* y_side, int idx) {
    // lat_is_y:  v = (c_row s_col, s_row, c_row c_col):  A_k = R[3k] s_col + R[3k+2] c_col (columns),  B_k = R[3k+1] s_row, scale c_row (rows)
    // otherwise: v = (s_col, c_col s_row, c_col c_row):  A_k = R[3k+1] s_row + R[3k+2] c_row (rows),   B_k = R[3k] s_col,  scale c_col (columns)
    const bool is_x = lat_is_y == is_row;
    if (is_x) {
        const int o = lat_is_y ? 1 : 0;
        x_side[idx] = cv;
        x_side[32 + idx] = R[o] * sv;
        x_side[64 + idx] = R[3 + o] * sv;
        x_side[96 + idx] = R[6 + o] * sv;
    } else {
        const int o = lat_is_y ? 0 : 1;
        y_side[idx] = fma(R[o], sv, R[2] * cv);
        y_side[32 + idx] = fma(R[3 + o], sv, R[5] * cv);
        y_side[64 + idx] = fma(R[6 + o], sv, R[8] * cv);
    }
}

// One pixel of the standard chain from its rotated unit vector v = scale A + B.  NORMALISED: stop before Denormalize
// (per-frame radius: the caller applies it).
template <bool NORMALISED>
__device__ __forceinline__ void std_pixel(double vx, double vy, double vz, const double* poly, int n_poly, const double* den,
                                          double& ox, double& oy) {
    const double h2 = fma(vx, vx, vy * vy);
    double ux = 0.0, uy = 1.0, h = h2;  // atan2(0, 0) = 0 -> (sin, cos) = (0, 1); a NaN h2 reaches theta
    if (h2 > 0.0) {
        const double inv = rsqrt(h2);
        ux = vx * inv;
        uy = vy * inv;
        h = h2 * inv;
    }
    double r = theta_unit(vz, h);
    if (r == 0.0) {  // PolarRollTransformer re-derives roll = atan2(0, 0) = 0 from (0, 0): (cos, sin) = (1, 0)
        ux = 1.0;
        uy = 0.0;
    }
    if (n_poly >= 0) {  // np.polyval(np.flip(coefs_reverse), theta); one uniform jump into the unrolled Horner chain
        static_assert(VR180_MAX_OP_PARAMS == 12, "extend the chain below");
        double acc = 0.0;
#define VR180_HORNER(i) acc = fma(acc, r, poly[i]);
        switch (n_poly) {
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
        r = acc;
    }
    // a negative radius flips (r, ux, uy) together and a zero one maps to the centre whatever the direction: the products
    // below are the same.  FisheyeDecoder("equidistant"): r = theta / (pi / 2)
    r *= kTwoOverPi;
    if (NORMALISED) {
        ox = r * ux;
        oy = r * uy;
    } else {
        ox = fma(r * den[0], ux, den[2]);
        oy = fma(r * den[1], uy, den[3]);
    }
}

}  // namespace vr180
