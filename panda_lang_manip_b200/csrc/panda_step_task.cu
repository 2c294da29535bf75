// panda_step_task.cu -- per-task instantiation of the step / reset / state kernels.  Compiled once per task with
// -DPG_TASK=<0..5> so the six translation units build in parallel (each holds the fully unrolled dynamics twice: f32, f64).
#include "panda_kernels.cuh"

#ifndef PG_TASK
#error "compile with -DPG_TASK=<task id>"
#endif

namespace pg {

extern long long g_launches;

template <typename T, int TASK> void launch_step(const EnvDev<T>& E, int ctrl, const StepIO& io, cudaStream_t st) {
    const int grid = ((E.perm ? E.tcount : E.n) + BLOCK - 1) / BLOCK;
    static bool configured = false;
    if (!configured) {      // > 48 KB of dynamic shared memory needs the opt-in
        cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes<T, TASK, CTRL_EE>());
        cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_JOINTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes<T, TASK, CTRL_JOINTS>());
        cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_EE>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(step_kernel<T, TASK, CTRL_JOINTS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured = true;
    }
    if (ctrl == CTRL_EE) step_kernel<T, TASK, CTRL_EE><<<grid, BLOCK, step_smem_bytes<T, TASK, CTRL_EE>(), st>>>(E, io);
    else step_kernel<T, TASK, CTRL_JOINTS><<<grid, BLOCK, step_smem_bytes<T, TASK, CTRL_JOINTS>(), st>>>(E, io);
    g_launches++;
}
template <typename T, int TASK> void launch_reset(const EnvDev<T>& E, const ResetIO& io, cudaStream_t st) {
    reset_kernel<T, TASK><<<(E.n + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(E, io);
    g_launches++;
}
template <typename T, int TASK> void launch_get_state(const EnvDev<T>& E, double* out, cudaStream_t st) {
    get_state_kernel<T, TASK><<<(E.n + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(E, out);
    g_launches++;
}
template <typename T, int TASK> void launch_set_state(const EnvDev<T>& E, const double* in, const unsigned char* mask, cudaStream_t st) {
    set_state_kernel<T, TASK><<<(E.n + BLOCK - 1) / BLOCK, BLOCK, 0, st>>>(E, in, mask);
    g_launches++;
}

#define PG_INST(T)                                                                                           \
    template void launch_step<T, PG_TASK>(const EnvDev<T>&, int, const StepIO&, cudaStream_t);               \
    template void launch_reset<T, PG_TASK>(const EnvDev<T>&, const ResetIO&, cudaStream_t);                  \
    template void launch_get_state<T, PG_TASK>(const EnvDev<T>&, double*, cudaStream_t);                     \
    template void launch_set_state<T, PG_TASK>(const EnvDev<T>&, const double*, const unsigned char*, cudaStream_t);
PG_INST(float)
PG_INST(double)

}  // namespace pg
