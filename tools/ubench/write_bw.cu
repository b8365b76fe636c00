// write_bw.cu — write-only HBM bandwidth of a B200 three ways: cudaMemsetAsync, a plain 16-byte-store kernel, and TMA bulk
// stores (cp.async.bulk.global.shared::cta) of 84480-byte blocks from shared memory, which is the access pattern of
// gen_vertices_kernel's output (one contiguous 660-line block per work item).  Usage: ./write_bw [GiB]   (prints GB/s each)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void store16(uint4* p, size_t n16) {
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void copy16(const uint4* __restrict__ s, uint4* __restrict__ d, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

constexpr uint32_t BLOCK_BYTES = 84480;
__global__ void __launch_bounds__(128, 1) tma_store(unsigned char* out, uint32_t n_blocks, int chunks) {
    extern __shared__ __align__(128) unsigned char img[];
    for (uint32_t i = threadIdx.x; i < BLOCK_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(img)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(img);
        for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
            unsigned char* dst = out + (size_t)b * BLOCK_BYTES;
            const uint32_t per = BLOCK_BYTES / chunks;
            for (int c = 0; c < chunks; ++c)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)c * per), "r"(sbase + c * per), "r"(per) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // two groups in flight, like the kernel's two images
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main(int argc, char** argv) {
    const double gib = argc > 1 ? atof(argv[1]) : 1.5;
    const size_t bytes = ((size_t)(gib * (1ull << 30)) / BLOCK_BYTES) * BLOCK_BYTES;
    unsigned char *d = nullptr, *s = nullptr;
    CK(cudaMalloc(&d, bytes));
    CK(cudaMalloc(&s, bytes));
    CK(cudaMemset(s, 1, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto report = [&](const char* name, float ms, double moved) { std::printf("%-34s %8.3f ms  %8.1f GB/s\n", name, ms, moved / (ms * 1e-3) / 1e9); };
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); CK(cudaMemsetAsync(d, 7, bytes)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("cudaMemsetAsync (write only)", ms, (double)bytes);
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); store16<<<sms * 8, 256>>>((uint4*)d, bytes / 16); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("st.global.v4 kernel (write only)", ms, (double)bytes);
    }
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); copy16<<<sms * 8, 256>>>((const uint4*)s, (uint4*)d, bytes / 16); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); if (rep) report("copy kernel (read + write bytes)", ms, 2.0 * bytes);
    }
    CK(cudaFuncSetAttribute(tma_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BLOCK_BYTES));
    const int chunk_counts[2] = {1, 16};
    for (int k = 0; k < 2; ++k) {
        const int chunks = chunk_counts[k];
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0); tma_store<<<sms, 128, BLOCK_BYTES>>>(d, (uint32_t)(bytes / BLOCK_BYTES), chunks); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            char name[64]; std::snprintf(name, sizeof(name), "TMA bulk store, %d x %u B / block", chunks, BLOCK_BYTES / chunks);
            if (rep) report(name, ms, (double)bytes);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
