// scratch: resource query of fused_rows_kernel on the device (nvcc -gencode arch=compute_100a,code=sm_100a fr_occ.cu -o fr_occ)
#include <cstdio>
#include "../recommender_tensorflow_b200/csrc/fused_rows.cu"
int main() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, fused_rows_kernel<16, 16, 2>);
    printf("regs %d static smem %zu local %zu maxThreads %d maxDyn %d\n", fa.numRegs, fa.sharedSizeBytes, fa.localSizeBytes, fa.maxThreadsPerBlock, fa.maxDynamicSharedSizeBytes);
    for (int smem = 200 * 1024; smem <= 232448; smem += 4096) {
        cudaError_t e = cudaFuncSetAttribute(fused_rows_kernel<16, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int nb = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fused_rows_kernel<16, 16, 2>, FR_THREADS, smem);
        printf("smem %d set=%d blocks/SM=%d\n", smem, (int)e, nb);
    }
    for (int th = 256; th <= 1024; th += 32) {
        int nb = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fused_rows_kernel<16, 16, 2>, th, 100 * 1024);
        printf("threads %d blocks/SM=%d\n", th, nb);
    }
    return 0;
}
