// does a kernel on a second stream run beside a persistent kernel that vacates some SMs?
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void persistent(volatile int* flag, int sm_limit, int* waited, int dyn)
{
    extern __shared__ double2 sm[];
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)smid >= sm_limit) return;
    if (dyn && threadIdx.x == 0) sm[0] = make_double2(1, 2);
    long long spins = 0;
    while (*flag == 0 && spins < (1 << 22)) { __nanosleep(500); spins++; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *waited = (int)(spins >> 4);
}
__global__ void setter(int* flag) { __threadfence(); *flag = 1; }
int main()
{
    int *flag, *waited;
    cudaMalloc(&flag, 4); cudaMalloc(&waited, 4);
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaFuncSetAttribute(persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // variant bit 2: stream a carries an L2 access-policy window, like the library's launch stream
    double* big; cudaMalloc(&big, 16 << 20);
    for (int variant = 0; variant < 8; variant++) {
        if (variant == 4) {
            cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 16 << 20);
            cudaStreamAttrValue attr; memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = big; attr.accessPolicyWindow.num_bytes = 16 << 20; attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            printf("set window: %d\n", (int)cudaStreamSetAttribute(a, cudaStreamAttributeAccessPolicyWindow, &attr));
        }
        const int sm_limit = (variant & 1) ? 140 : 100000, dyn = (variant & 2) ? 9 * 1024 : 0;
        cudaMemset(flag, 0, 4); cudaMemset(waited, 0, 4);
        cudaEvent_t e; cudaEventCreate(&e);
        cudaEventRecord(e, a);
        persistent<<<148 * 4, 128, dyn, a>>>(flag, sm_limit, waited, dyn);
        cudaStreamWaitEvent(b, e, 0);
        setter<<<1, 1, 0, b>>>(flag);
        cudaDeviceSynchronize();
        int w; cudaMemcpy(&w, waited, 4, cudaMemcpyDeviceToHost);
        printf("sm_limit %d dyn_smem %d: block 0 waited %d x16 spins (%s)\n", sm_limit, dyn, w, w < (1 << 17) ? "setter ran beside" : "setter did NOT run beside");
    }
    return 0;
}
