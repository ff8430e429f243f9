// Fused tail of the JPEG-transcode chain: dequantise + 8x8 inverse DCT (+ inverse YCbCr + clamp) in ONE launch for all
// components, instead of one k_quantize launch per coefficient plane (192 at 4096^2), one k_inv_dct per component and k_ycbcr.
//
// Reference semantics (bit-exact): transform/quantize.h:32-49 (v *= q with the int16 wrap of pixel_type), transform/dct.h:60-107,
// 249-296 (coefficient i of a block comes from plane ordering[c][zigzag[i]]; the DC gets (maxval+1)*4 added in float; columns,
// then rows; out = 0.0, out += k*in for u = 0..7 in IEEE double, no contraction -- the reference is built for baseline x86-64;
// round() half away from zero, then the narrowing to int16), transform/ycbcr.h:49-58 (float loads, double arithmetic in source
// order, +0.5, CLAMP, truncation), image/image.cpp:107-113 (final clamp: the identity after the YCbCr clamp).
//
// Work split: a CTA of 8 warps owns 32 consecutive 8x8 blocks of one block row (a 256 x 8 strip of pixels).  Column pass: warp c,
// lane b computes column c of block b (its 8 coefficients are 8 coalesced 64-byte reads of 8 coefficient planes); the strip of
// 32 x 64 doubles is exchanged through shared memory in [row][column][block] order (conflict-free both ways); row pass: warp y,
// lane b computes row y of block b.  Products of equal magnitude are computed once: (-k)*in == -(k*in) and a + (-p) == a - p
// exactly in IEEE arithmetic, so the results are bit-identical to the reference's 64 multiplications per 1-D transform with 22.
// The three components' rounded samples meet in shared memory, and the colour inverse writes R, G, B with 16-byte stores.
// FP64-bound by design: ~2 x 78 double operations per 8 samples, ~60 us of DFMA-pipe time at 4096^2 x 3 on B200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace idf {

struct Job {
    const int16_t *pl[3][64];   // coefficient planes of each component in block order (row-major 8x8); nullptr = all zero
    int q[3][64];               // dequantisation factor of each plane (1 = none)
    int16_t *out[3];            // output planes, 8*bw x 8*bh
    int ncomp, bw, bh;
    float dc_offset;
    int ycbcr;                  // 1: inverse YCbCr + clamp on the three components before they are stored
    int minval, maxval;
};

#define IDF_C0 0.3535533906
#define IDF_C1 0.4903926402
#define IDF_C2 0.4619397663
#define IDF_C3 0.4157348062
#define IDF_C5 0.2777851165
#define IDF_C6 0.1913417162
#define IDF_C7 0.0975451610

__device__ __forceinline__ double dm(double k, double x) { return __dmul_rn(k, x); }
__device__ __forceinline__ double da(double a, double p) { return __dadd_rn(a, p); }

// out[x] = sum_u kDCTMatrix[8u + x] * in[u], accumulated from 0.0 in the order u = 0..7 (IDCT1d, dct.h:88-95)
__device__ __forceinline__ void idct8(const double (&in)[8], double (&o)[8]) {
    double p, q, r, s;
    p = dm(IDF_C0, in[0]);
#pragma unroll
    for (int x = 0; x < 8; x++) o[x] = da(0.0, p);
    p = dm(IDF_C1, in[1]); q = dm(IDF_C3, in[1]); r = dm(IDF_C5, in[1]); s = dm(IDF_C7, in[1]);
    o[0] = da(o[0], p); o[1] = da(o[1], q); o[2] = da(o[2], r); o[3] = da(o[3], s); o[4] = da(o[4], -s); o[5] = da(o[5], -r); o[6] = da(o[6], -q); o[7] = da(o[7], -p);
    p = dm(IDF_C2, in[2]); q = dm(IDF_C6, in[2]);
    o[0] = da(o[0], p); o[1] = da(o[1], q); o[2] = da(o[2], -q); o[3] = da(o[3], -p); o[4] = da(o[4], -p); o[5] = da(o[5], -q); o[6] = da(o[6], q); o[7] = da(o[7], p);
    p = dm(IDF_C3, in[3]); q = dm(IDF_C7, in[3]); r = dm(IDF_C1, in[3]); s = dm(IDF_C5, in[3]);
    o[0] = da(o[0], p); o[1] = da(o[1], -q); o[2] = da(o[2], -r); o[3] = da(o[3], -s); o[4] = da(o[4], s); o[5] = da(o[5], r); o[6] = da(o[6], q); o[7] = da(o[7], -p);
    p = dm(IDF_C0, in[4]);
    o[0] = da(o[0], p); o[1] = da(o[1], -p); o[2] = da(o[2], -p); o[3] = da(o[3], p); o[4] = da(o[4], p); o[5] = da(o[5], -p); o[6] = da(o[6], -p); o[7] = da(o[7], p);
    p = dm(IDF_C5, in[5]); q = dm(IDF_C1, in[5]); r = dm(IDF_C7, in[5]); s = dm(IDF_C3, in[5]);
    o[0] = da(o[0], p); o[1] = da(o[1], -q); o[2] = da(o[2], r); o[3] = da(o[3], s); o[4] = da(o[4], -s); o[5] = da(o[5], -r); o[6] = da(o[6], q); o[7] = da(o[7], -p);
    p = dm(IDF_C6, in[6]); q = dm(IDF_C2, in[6]);
    o[0] = da(o[0], p); o[1] = da(o[1], -q); o[2] = da(o[2], q); o[3] = da(o[3], -p); o[4] = da(o[4], -p); o[5] = da(o[5], q); o[6] = da(o[6], -q); o[7] = da(o[7], p);
    p = dm(IDF_C7, in[7]); q = dm(IDF_C5, in[7]); r = dm(IDF_C3, in[7]); s = dm(IDF_C1, in[7]);
    o[0] = da(o[0], p); o[1] = da(o[1], -q); o[2] = da(o[2], r); o[3] = da(o[3], -s); o[4] = da(o[4], s); o[5] = da(o[5], -r); o[6] = da(o[6], q); o[7] = da(o[7], -p);
}

__device__ __forceinline__ int ycc_clamp_trunc(double x, int lo, int hi) {     // CLAMP in double, then the truncating conversion (ycbcr.h:55-57)
    const double v = (x < (double)lo) ? (double)lo : ((x > (double)hi) ? (double)hi : x);
    return (int)(short)__double2int_rz(v);
}

constexpr int kBlocksPerCta = 32;
constexpr int kPixPitch = kBlocksPerCta * 8 + 8;        // halfwords per strip row in shared memory (16-byte aligned rows)

__global__ void __launch_bounds__(256) k_idct_ycbcr(const __grid_constant__ Job J) {
    __shared__ double tmp_s[64 * kBlocksPerCta];                // [y * 8 + c][b]
    __shared__ __align__(16) int16_t pix_s[3][8][kPixPitch];    // rounded samples of the strip, per component
    const int w8 = (int)(threadIdx.x >> 5), b = (int)(threadIdx.x & 31);
    const int bx = (int)blockIdx.x * kBlocksPerCta + b, by = (int)blockIdx.y;
    const bool valid = bx < J.bw;
    const size_t idx = (size_t)by * J.bw + (valid ? bx : 0);
    for (int comp = 0; comp < J.ncomp; comp++) {
        double in[8], o[8];
        // ---- column w8 of block b: coefficients (u, w8), u = 0..7, dequantised with the int16 wrap
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int k = u * 8 + w8;
            const int16_t *p = J.pl[comp][k];
            int v = (valid && p) ? (int)p[idx] : 0;
            v = (int)(short)(v * J.q[comp][k]);
            in[u] = (k == 0) ? (double)__fadd_rn((float)v, J.dc_offset) : (double)v;
        }
        idct8(in, o);
#pragma unroll
        for (int y = 0; y < 8; y++) tmp_s[(y * 8 + w8) * kBlocksPerCta + b] = o[y];
        __syncthreads();
        // ---- row w8 of block b
#pragma unroll
        for (int u = 0; u < 8; u++) in[u] = tmp_s[(w8 * 8 + u) * kBlocksPerCta + b];
        idct8(in, o);
        uint32_t pk[4];
#pragma unroll
        for (int x = 0; x < 8; x += 2) {
            const int v0 = (int)(short)__double2int_rz(round(o[x])), v1 = (int)(short)__double2int_rz(round(o[x + 1]));
            pk[x >> 1] = (uint32_t)(uint16_t)v0 | ((uint32_t)(uint16_t)v1 << 16);
        }
        *reinterpret_cast<uint4 *>(&pix_s[comp][w8][b * 8]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        __syncthreads();
    }
    // ---- the strip leaves: thread (row w8, 8 pixels at b * 8), through the colour inverse when asked for
    const int ow = J.bw * 8;
    const int px = (int)blockIdx.x * kBlocksPerCta * 8 + b * 8;
    if (px >= ow) return;
    const size_t o = (size_t)(by * 8 + w8) * ow + px;
    if (!J.ycbcr) {
        for (int comp = 0; comp < J.ncomp; comp++) *reinterpret_cast<uint4 *>(J.out[comp] + o) = *reinterpret_cast<const uint4 *>(&pix_s[comp][w8][b * 8]);
        return;
    }
    const uint4 a0 = *reinterpret_cast<const uint4 *>(&pix_s[0][w8][b * 8]), a1 = *reinterpret_cast<const uint4 *>(&pix_s[1][w8][b * 8]),
                a2 = *reinterpret_cast<const uint4 *>(&pix_s[2][w8][b * 8]);
    const uint32_t w0[4] = {a0.x, a0.y, a0.z, a0.w}, w1[4] = {a1.x, a1.y, a1.z, a1.w}, w2[4] = {a2.x, a2.y, a2.z, a2.w};
    uint32_t r4[4], g4[4], b4[4];
    const float half = (float)((J.maxval + 1) / 2);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        r4[i] = 0; g4[i] = 0; b4[i] = 0;
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
            const float yy = (float)(int)(short)(w0[i] >> (16 * hlf));
            const float cb = __fsub_rn((float)(int)(short)(w1[i] >> (16 * hlf)), half);
            const float cr = __fsub_rn((float)(int)(short)(w2[i] >> (16 * hlf)), half);
            const double dy = (double)yy, dcb = (double)cb, dcr = (double)cr;
            const double r = __dadd_rn(__dadd_rn(dy, __dmul_rn(1.402, dcr)), 0.5);
            const double g = __dadd_rn(__dsub_rn(__dsub_rn(dy, __dmul_rn(0.344136, dcb)), __dmul_rn(0.714136, dcr)), 0.5);
            const double bl = __dadd_rn(__dadd_rn(dy, __dmul_rn(1.772, dcb)), 0.5);
            r4[i] |= (uint32_t)(uint16_t)ycc_clamp_trunc(r, J.minval, J.maxval) << (16 * hlf);
            g4[i] |= (uint32_t)(uint16_t)ycc_clamp_trunc(g, J.minval, J.maxval) << (16 * hlf);
            b4[i] |= (uint32_t)(uint16_t)ycc_clamp_trunc(bl, J.minval, J.maxval) << (16 * hlf);
        }
    }
    *reinterpret_cast<uint4 *>(J.out[0] + o) = make_uint4(r4[0], r4[1], r4[2], r4[3]);
    *reinterpret_cast<uint4 *>(J.out[1] + o) = make_uint4(g4[0], g4[1], g4[2], g4[3]);
    *reinterpret_cast<uint4 *>(J.out[2] + o) = make_uint4(b4[0], b4[1], b4[2], b4[3]);
}

}  // namespace idf
