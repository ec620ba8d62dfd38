// Micro-benchmark: how long does the issuing thread spend per cp.async.bulk (UBLKCP), as a function of copy size and of the
// number of copies in flight?  One thread per CTA issues COPIES bulk copies back to back into a ring of shared-memory slots
// and records clock64 around every issue; a second kernel variant issues from four different warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_issue tools/micro/tma_issue.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
static __device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void __launch_bounds__(128) k(const char* __restrict__ src, size_t per_cta, int copies, int bytes, int slots, long long* out, int nwarp_issue) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar[16];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 16; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[i])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const char* base = src + static_cast<size_t>(blockIdx.x) * per_cta;
    const int warp = tid >> 5, lane = tid & 31;
    long long t_issue = 0, t_wait = 0;
    if (lane == 0 && warp < nwarp_issue) {
        uint32_t phase[16] = {0};
        for (int c = warp; c < copies; c += nwarp_issue) {
            const int s = c % slots;
            if (c >= slots) {   // the slot's previous copy must have landed
                const long long w0 = clock64();
                asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&bar[s])), "r"(phase[s]) : "memory");
                phase[s] ^= 1;
                t_wait += clock64() - w0;
            }
            const long long t0 = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + s * bytes)),
                         "l"(base + static_cast<size_t>(c) * bytes), "r"(bytes), "r"(s32(&bar[s]))
                         : "memory");
            t_issue += clock64() - t0;
        }
        // drain
        for (int s = 0; s < slots && s < copies; s++) {
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&bar[s])), "r"(phase[s]) : "memory");
        }
        atomicAdd(reinterpret_cast<unsigned long long*>(out), static_cast<unsigned long long>(t_issue));
        atomicAdd(reinterpret_cast<unsigned long long*>(out + 1), static_cast<unsigned long long>(t_wait));
    }
}
int main() {
    const int ctas_per_sm[] = {1, 3};
    const int sizes[] = {1024, 2048, 8192, 16384};
    const int slotsv[] = {1, 3, 12};
    char* src;
    const size_t per_cta = 4u << 20;
    const int max_ctas = 148 * 3;
    cudaMalloc(&src, per_cta * max_ctas);
    cudaMemset(src, 1, per_cta * max_ctas);
    long long* out;
    cudaMalloc(&out, 16);
    for (int nw : {1})   // (one issuing thread per CTA; slots are not shared between issuers)
    for (int cps : ctas_per_sm)
        for (int bytes : sizes)
            for (int slots : slotsv) {
                if (slots > 12 || slots * bytes > 70000) continue;
                const int copies = (2 << 20) / bytes;   // 2 MB per CTA
                const int grid = 148 * cps;
                cudaMemset(out, 0, 16);
                const size_t smem = 72 * 1024;   // forces <= 3 CTAs per SM
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                k<<<grid, 128, smem>>>(src, per_cta, copies, bytes, slots, out, nw);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("issuing warps %d  CTAs/SM %d  copy %5d B  slots %2d : %7.1f cycles per issue, %7.1f cycles waiting per copy, %6.1f GB/s  (%s)\n", nw, cps, bytes, slots,
                       (double)h[0] / ((double)copies * grid), (double)h[1] / ((double)copies * grid), (double)copies * bytes * grid / ms / 1e6, cudaGetErrorString(e));
            }
    return 0;
}
