// TMA tile-movement microbenchmark (sm_100a): how fast can independent warps stream narrow 2-D boxes of an int16 plane
// through shared memory?  Development aid for the packed horizontal unsqueeze kernel (fb_pk_squeeze.cuh), whose lanes own
// rows and therefore want tall, narrow tiles.  Each warp loads K boxes (2-stage ring, mbarrier), optionally stores them back.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tma ubench_tma.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, int bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, int parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tile_load(void *dst, const CUtensorMap *m, int x, int y, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)), "l"(m), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tile_store(const CUtensorMap *m, int x, int y, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(m), "r"(x), "r"(y), "r"(smem_u32(src)) : "memory");
}

// every warp: tiles (tx0 + k, ty) for k = 0..K-1 of the source map; optional store to the destination map
__global__ void k_stream(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst, int box_w, int box_h, int K, int tiles_x, int tiles_y,
                         int do_store, int stages, unsigned *sink) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int tile_bytes = box_w * box_h * 2;
    unsigned char *sm = smraw + (size_t)warp * (stages * tile_bytes + 128);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + stages * tile_bytes);
    if (lane == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int gw = blockIdx.x * wpb + warp;
    const int runs_x = tiles_x / K;
    const int ty = (gw / runs_x) % tiles_y, tx0 = (gw % runs_x) * K;
    unsigned acc = 0;
    int par = 0;
    if (lane == 0)
        for (int c = 0; c < stages - 1 && c < K; c++) { mbar_expect(&bars[c], tile_bytes); tile_load(sm + c * tile_bytes, &src, (tx0 + c) * box_w, ty * box_h, &bars[c]); }
    for (int c = 0; c < K; c++) {
        const int st = c % stages;
        __syncwarp();
        if (lane == 0 && c + stages - 1 < K) {
            const int s2 = (c + stages - 1) % stages;
            if (do_store) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_expect(&bars[s2], tile_bytes);
            tile_load(sm + s2 * tile_bytes, &src, (tx0 + c + stages - 1) * box_w, ty * box_h, &bars[s2]);
        }
        mbar_wait(&bars[st], (par >> st) & 1);
        par ^= 1 << st;
        acc += *reinterpret_cast<const unsigned *>(sm + st * tile_bytes + lane * 4);
        if (do_store) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) { tile_store(&dst, (tx0 + c) * box_w, ty * box_h, sm + st * tile_bytes); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
        }
    }
    if (lane == 0 && do_store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 0x12345678u) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeFn enc = (EncodeFn)fnp;
    const int W = 4096, H = 4096;
    int16_t *a, *b;
    unsigned *sink;
    cudaMalloc(&a, (size_t)W * H * 2); cudaMalloc(&b, (size_t)W * H * 2); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, (size_t)W * H * 2);
    char *flush; cudaMalloc(&flush, 256 << 20);
    const int shapes[][2] = {{16, 64}, {24, 64}, {32, 64}, {64, 64}, {32, 32}, {64, 32}, {128, 32}, {128, 16}, {256, 8}};
    for (auto &sh : shapes) {
        const int bw = sh[0], bh = sh[1];
        CUtensorMap ms, md;
        const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W * 2};
        const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
        for (int promo = 0; promo < 2; promo++) {
            const CUtensorMapL2promotion pr = promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
            enc(&ms, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            enc(&md, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, b, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            for (int do_store = 0; do_store < 2; do_store++) {
                for (int wps : {4, 8, 12}) {
                    const int stages = 2, K = 16;
                    const int tiles_x = W / bw, tiles_y = H / bh;
                    const int total_warps = (tiles_x / K) * tiles_y;            // covers the plane exactly once
                    const int wpb = 1;
                    const size_t smem = (size_t)wpb * (stages * bw * bh * 2 + 128);
                    if (smem * wps > 220 * 1024) continue;
                    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem));
                    // occupancy is capped through shared memory: pad the request so that exactly wps blocks fit one SM
                    const size_t pad = (220 * 1024) / wps;
                    const size_t req = pad > smem ? pad : smem;
                    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)req);
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0); cudaEventCreate(&e1);
                    float best = 1e9f;
                    for (int rep = 0; rep < 4; rep++) {
                        cudaMemset(flush, rep, 256 << 20);
                        cudaEventRecord(e0);
                        k_stream<<<total_warps / wpb, 32 * wpb, req>>>(ms, md, bw, bh, K, tiles_x, tiles_y, do_store, stages, sink);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms_ = 0;
                        cudaEventElapsedTime(&ms_, e0, e1);
                        if (rep > 0 && ms_ < best) best = ms_;
                    }
                    const double bytes = (double)W * H * 2 * (do_store ? 2 : 1);
                    printf("box %3dx%-3d (%3d B rows) promo128=%d %s warps/SM=%2d : %7.1f us  %7.1f GB/s   err=%s\n", bw, bh, bw * 2, promo, do_store ? "load+store" : "load      ", wps,
                           best * 1000.f, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
                }
            }
        }
    }
    return 0;
}
