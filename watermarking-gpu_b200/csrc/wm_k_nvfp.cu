// wm_k_nvfp.cu — NVF mask planes for the window sizes p = 5, 7, 9 (see k_nvfp in wm_kernels.cuh)
#include "wm_launch.h"

namespace wm {

template <typename PixT, int PW>
static void launch_nvfp_t(bool tr, dim3 grid, cudaStream_t st, const NvfpArgs& a)
{
    if (tr) k_nvfp<PixT, PW, true><<<grid, NT, 0, st>>>(a);
    else k_nvfp<PixT, PW, false><<<grid, NT, 0, st>>>(a);
}
template <typename PixT>
static void launch_nvfp_p(int pw, bool tr, dim3 grid, cudaStream_t st, const NvfpArgs& a)
{
    if (pw == 5) launch_nvfp_t<PixT, 5>(tr, grid, st, a);
    else if (pw == 7) launch_nvfp_t<PixT, 7>(tr, grid, st, a);
    else launch_nvfp_t<PixT, 9>(tr, grid, st, a);
}
void launch_nvfp(int dtype, int pw, bool tr, dim3 grid, cudaStream_t st, const NvfpArgs& a)
{
    if (dtype == WM_F32) launch_nvfp_p<float>(pw, tr, grid, st, a);
    else launch_nvfp_p<uint8_t>(pw, tr, grid, st, a);
}

}  // namespace wm
