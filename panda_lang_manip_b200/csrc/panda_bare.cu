// panda_bare.cu -- instantiation of the bare-world step (the sim facade used without a task: reference panda_gym/pybullet.py
// loadURDF / create_box / control_joints / step, exercised by the reference's own test/pybullet_test.py).
#include "panda_kernels.cuh"

namespace pg {

extern long long g_launches;

template <typename T, int NOBJ> static cudaError_t configure_one(void) {
    return cudaFuncSetAttribute(bare_step_kernel<T, NOBJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bare_smem_bytes<T, NOBJ>());
}
template <typename T> cudaError_t configure_bare(void) {
    cudaError_t e;
    if ((e = configure_one<T, 0>()) != cudaSuccess) return e;
    if ((e = configure_one<T, 1>()) != cudaSuccess) return e;
    return configure_one<T, 2>();
}
template <typename T> void launch_bare_step(const EnvDev<T>& E, int nobj, int nsub, cudaStream_t st) {
    const int grid = (E.n + BARE_BLOCK - 1) / BARE_BLOCK;
    if (nobj == 0) bare_step_kernel<T, 0><<<grid, BARE_BLOCK, bare_smem_bytes<T, 0>(), st>>>(E, nsub);
    else if (nobj == 1) bare_step_kernel<T, 1><<<grid, BARE_BLOCK, bare_smem_bytes<T, 1>(), st>>>(E, nsub);
    else bare_step_kernel<T, 2><<<grid, BARE_BLOCK, bare_smem_bytes<T, 2>(), st>>>(E, nsub);
    g_launches++;
}
template void launch_bare_step<float>(const EnvDev<float>&, int, int, cudaStream_t);
template void launch_bare_step<double>(const EnvDev<double>&, int, int, cudaStream_t);
template cudaError_t configure_bare<float>(void);
template cudaError_t configure_bare<double>(void);

}  // namespace pg
