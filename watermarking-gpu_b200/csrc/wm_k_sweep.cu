// wm_k_sweep.cu — instantiations + dispatch of one kernel family (see wm_launch.h)
#include "wm_launch.h"

namespace wm {

void launch_sweep(int dtype, bool fp16, bool tma, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const SweepArgs& a)
{
    if (dtype == WM_F32) {
        if (tma) { if (fp16) WM_LAUNCH((k_sweep<float, true, true>), sweep_smem(true, false), tmI, a); else WM_LAUNCH((k_sweep<float, false, true>), sweep_smem(true, false), tmI, a); }
        else { if (fp16) WM_LAUNCH((k_sweep<float, true, false>), sweep_smem(false, false), tmI, a); else WM_LAUNCH((k_sweep<float, false, false>), sweep_smem(false, false), tmI, a); }
    } else {
        if (tma) { if (fp16) WM_LAUNCH((k_sweep<uint8_t, true, true>), sweep_smem(true, true), tmI, a); else WM_LAUNCH((k_sweep<uint8_t, false, true>), sweep_smem(true, true), tmI, a); }
        else { if (fp16) WM_LAUNCH((k_sweep<uint8_t, true, false>), sweep_smem(false, true), tmI, a); else WM_LAUNCH((k_sweep<uint8_t, false, false>), sweep_smem(false, true), tmI, a); }
    }
}

}  // namespace wm
