// wm_k_stats.cu — instantiations + dispatch of one kernel family (see wm_launch.h)
#include "wm_launch.h"

namespace wm {

template <typename PixT, bool TMA>
void launch_stats_t(int mask, bool tr, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const CUtensorMap& tmW, const EmbedArgs& a)
{
    if (mask == 2) { WM_LAUNCH((k_stats<PixT, 2, false, TMA>), embed_smem(TMA, sizeof(PixT) == 1), tmI, tmW, a); }  // NVF plane (p > 3)
    else if (mask == WM_MASK_ME) { if (tr) WM_LAUNCH((k_stats<PixT, 0, true, TMA>), embed_smem(TMA, sizeof(PixT) == 1), tmI, tmW, a); else WM_LAUNCH((k_stats<PixT, 0, false, TMA>), embed_smem(TMA, sizeof(PixT) == 1), tmI, tmW, a); }
    else { if (tr) WM_LAUNCH((k_stats<PixT, 1, true, TMA>), embed_smem(TMA, sizeof(PixT) == 1), tmI, tmW, a); else WM_LAUNCH((k_stats<PixT, 1, false, TMA>), embed_smem(TMA, sizeof(PixT) == 1), tmI, tmW, a); }
}
// u8 TMA frames on 128-thread CTAs (8 lines per thread, 4 CTAs per SM)
static void launch_stats_u8n(int mask, bool tr, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const CUtensorMap& tmW, const EmbedArgs& a)
{
    constexpr int SM = embed_smem(true, true, ENT_U8);
    if (mask == 2) { WM_LAUNCH_T((k_stats<uint8_t, 2, false, true, ENT_U8>), ENT_U8, SM, tmI, tmW, a); }
    else if (mask == WM_MASK_ME) { if (tr) WM_LAUNCH_T((k_stats<uint8_t, 0, true, true, ENT_U8>), ENT_U8, SM, tmI, tmW, a); else WM_LAUNCH_T((k_stats<uint8_t, 0, false, true, ENT_U8>), ENT_U8, SM, tmI, tmW, a); }
    else { if (tr) WM_LAUNCH_T((k_stats<uint8_t, 1, true, true, ENT_U8>), ENT_U8, SM, tmI, tmW, a); else WM_LAUNCH_T((k_stats<uint8_t, 1, false, true, ENT_U8>), ENT_U8, SM, tmI, tmW, a); }
}
void launch_stats(int dtype, int mask, bool tr, bool tma, bool narrow, dim3 grid, cudaStream_t st, const CUtensorMap& tmI, const CUtensorMap& tmW, const EmbedArgs& a)
{
    if (narrow && tma && dtype != WM_F32) { launch_stats_u8n(mask, tr, grid, st, tmI, tmW, a); return; }
    if (dtype == WM_F32) { if (tma) launch_stats_t<float, true>(mask, tr, grid, st, tmI, tmW, a); else launch_stats_t<float, false>(mask, tr, grid, st, tmI, tmW, a); }
    else { if (tma) launch_stats_t<uint8_t, true>(mask, tr, grid, st, tmI, tmW, a); else launch_stats_t<uint8_t, false>(mask, tr, grid, st, tmI, tmW, a); }
}

}  // namespace wm
