// wm_k_sweep.cu — instantiations + dispatch of one kernel family (see wm_launch.h)
#include "wm_launch.h"

namespace wm {

bool& pdl_next()
{
    static thread_local bool v = false;
    return v;
}

template <typename PixT, bool TMA>
static void launch_sweep_t(int acc, int smem, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const SweepArgs& a)
{
    if (acc == 2) WM_LAUNCH_T((k_sweep<PixT, 2, TMA>), SNT, smem, tmI, a);
    else if (acc == 1) WM_LAUNCH_T((k_sweep<PixT, 1, TMA>), SNT, smem, tmI, a);
    else WM_LAUNCH_T((k_sweep<PixT, 0, TMA>), SNT, smem, tmI, a);
}

// acc: 0 = f32 products, 1 = fp16-rounded products + FHADD accumulation, 2 = fp16-rounded products + HMMA accumulation
void launch_sweep(int dtype, int acc, bool tma, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const SweepArgs& a)
{
    if (dtype == WM_F32) {
        if (tma) launch_sweep_t<float, true>(acc, sweep_smem(true, false), grid, st, tmI, a);
        else launch_sweep_t<float, false>(acc, sweep_smem(false, false), grid, st, tmI, a);
    } else {
        if (tma) launch_sweep_t<uint8_t, true>(acc, sweep_smem(true, true), grid, st, tmI, a);
        else launch_sweep_t<uint8_t, false>(acc, sweep_smem(false, true), grid, st, tmI, a);
    }
}

}  // namespace wm
