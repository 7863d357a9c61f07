// Driver: loads the two cubins (ptxas -O1 / -O3) and times every variant with CUDA events on 148 CTAs x 256 threads.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <string>
#include <vector>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s; cuGetErrorString(r_, &s); printf("%s failed: %s\n", #x, s); return 1; } } while (0)
int main(int argc, char** argv) {
    const char* dir = argc > 1 ? argv[1] : ".";
    cudaFree(0);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = sizeof(float) * (8 * 10 * 260 + 72 * 256);
    for (const char* sfx : {"O1", "O3"}) {
        CUmodule mod;
        CK(cuModuleLoad(&mod, (std::string(dir) + "/ffma2_kernels_" + sfx + ".cubin").c_str()));
        for (const char* base : {"k_loop0_", "k_loop1_", "k_loop2_", "k_reg1_", "k_reg0_"}) {
            CUfunction f;
            std::string name = std::string(base) + sfx;
            CK(cuModuleGetFunction(&f, mod, name.c_str()));
            const bool loop = name.find("loop") != std::string::npos;
            if (loop) CK(cuFuncSetAttribute(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
            int iters = loop ? 400 : 20000;
            float seed = 1.f;
            void* args_loop[] = {&out, &iters};
            void* args_reg[] = {&out, &iters, &seed};
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0, 0);
                CK(cuLaunchKernel(f, sms, 1, 1, 256, 1, 1, loop ? (unsigned)smem : 0, 0, loop ? args_loop : args_reg, nullptr));
                cudaEventRecord(e1, 0);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const double fma = loop ? (double)iters * 256 * 40 * 2 * 256.0 * sms : (double)iters * 64 * 2 * 256.0 * sms;
            printf("%-14s %8.3f ms  %7.2f TFLOP/s\n", name.c_str(), best, 2.0 * fma / (best * 1e-3) / 1e12);
        }
    }
    return 0;
}
