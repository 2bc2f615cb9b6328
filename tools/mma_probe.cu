// Microbenchmark: cycles per tcgen05.mma kind::tf32 (M=128, K=8, SS operands, K-major SWIZZLE_128B) as a
// function of N and of how many MMAs are issued per commit.  One CTA per SM, operands are whatever is in smem.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I heat_b200/csrc -I include tools/mma_probe.cu -o gpurun_out/mma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "hk_tma.cuh"
using namespace hk;

__global__ void __launch_bounds__(128, 1) probe(int N, int per_commit, int reps, int lag, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tslot;
    __shared__ uint64_t bar[8];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    if (warp == 1) {
        const uint32_t sb = smem_u32(smem);
        const uint32_t idesc = umma_idesc_tf32(128, N);
        const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
        // round r commits to barrier r % 8 and then waits for round r - lag (lag = 0: full round trip)
        long long tiss = 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const long long ta = clock64();
            if (elect_one()) {
                for (int m = 0; m < per_commit; ++m) {
                    const uint64_t ad = umma_desc_k_sw128(sb + (m & 3) * 32);
                    const uint64_t bd = umma_desc_k_sw128(sb + 16384 + (m & 3) * 32);
                    umma_tf32(tb + (uint32_t)((r & 1) * N), ad, bd, idesc, m ? 1u : 0u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[r & 7])) : "memory");
            }
            __syncwarp();
            tiss += clock64() - ta;
            const int w = r - lag;
            if (w >= 0) mbar_wait_a(smem_u32(&bar[w & 7]), (uint32_t)((w >> 3) & 1));
        }
        for (int w = reps - lag; w < reps; ++w)
            if (w >= 0) mbar_wait_a(smem_u32(&bar[w & 7]), (uint32_t)((w >> 3) & 1));
        long long t1 = clock64();
        if (threadIdx.x == 32 && blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = tiss;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tbase, 512);
    }
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int reps = 2000;
    for (int lag : {0, 1, 2, 4})
        for (int N : {64, 128})
            for (int pc : {4, 5, 8}) {
                probe<<<148, 128, 64 * 1024>>>(N, pc, reps, lag, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2] = {0, 0};
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("lag=%d N=%3d mma/commit=%2d : %7.1f cycles per round, %6.1f per mma, issue section %6.1f  (%s)\n", lag, N, pc,
                       (double)h[0] / reps, (double)h[0] / reps / pc, (double)h[1] / reps, cudaGetErrorString(e));
            }
    return 0;
}
