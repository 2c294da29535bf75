// panda_step_task.cu -- per-task instantiation of the step / reset / state kernels.  Compiled once per task with
// -DPG_TASK=<0..5> so the six translation units build in parallel (each holds the fully unrolled dynamics twice: f32, f64).
#include "panda_kernels.cuh"

#ifndef PG_TASK
#error "compile with -DPG_TASK=<task id>"
#endif

namespace pg {

extern long long g_launches;

template <typename T, int TASK> void launch_step(const EnvDev<T>& E, int ctrl, const StepIO& io, cudaStream_t st) {
    constexpr int BS = step_block<T, TASK>();
    const int grid = ((E.perm ? E.tcount : E.n) + BS - 1) / BS;
    if (ctrl == CTRL_EE) step_kernel<T, TASK, CTRL_EE><<<grid, BS, step_smem_bytes<T, TASK, CTRL_EE>(), st>>>(E, io);
    else step_kernel<T, TASK, CTRL_JOINTS><<<grid, BS, step_smem_bytes<T, TASK, CTRL_JOINTS>(), st>>>(E, io);
    g_launches++;
}
// > 48 KB of dynamic shared memory needs an opt-in per function AND per device (function attributes live in the device's context):
// pg_create calls this for the handle's device; errors are reported to the caller.
template <typename T, int TASK> cudaError_t configure_step(void) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes<T, TASK, CTRL_EE>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_JOINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes<T, TASK, CTRL_JOINTS>())) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_EE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_JOINTS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
template <typename T, int TASK> void launch_reset(const EnvDev<T>& E, const ResetIO& io, cudaStream_t st) {
    reset_kernel<T, TASK><<<(E.n + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(E, io);
    g_launches++;
}
#define PG_INST(T)                                                                                           \
    template void launch_step<T, PG_TASK>(const EnvDev<T>&, int, const StepIO&, cudaStream_t);               \
    template void launch_reset<T, PG_TASK>(const EnvDev<T>&, const ResetIO&, cudaStream_t);                  \
    template cudaError_t configure_step<T, PG_TASK>(void);
PG_INST(float)
PG_INST(double)

}  // namespace pg
