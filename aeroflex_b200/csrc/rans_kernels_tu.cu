// rans_kernels_tu.cu -- one translation unit per arithmetic mode:
//   nvcc -DAFX_FAST=0 -fmad=false -> namespace afx::strict (bit-identical to the CPU reference)
//   nvcc -DAFX_FAST=1 -fmad=true  -> namespace afx::fast   (shared reciprocals, FMA contraction)
#include "rans_kernels.cuh"
#include "rans_krylov.cuh"
#include "rans_stage.cuh"
#include "rans_pipe.cuh"

namespace afx {
namespace AFX_NS {

const KernelTable& table()
{
    static const KernelTable t = {
#if AFX_FAST
        "fast",
#else
        "strict",
#endif
        launch::dt_grad, launch::limiter, launch::flux, launch::gather, launch::stage, launch::pipe, launch::pipe_item_elems, launch::pipe_ctas_per_sm, launch::stage_prepare, launch::stage_threads, launch::tile_k3a, launch::face_record_kinds, launch::tile_face_kinds, launch::halo_signal, launch::halo_wait_scatter, launch::gather_blocks, launch::norm_finish,
        launch::jacobian, launch::jac_diag,
        launch::wall_forces, launch::prolongate, launch::fill_cells, launch::ghost_fill, launch::ghost_follow, launch::permute4, launch::permute1, launch::scatter4,
        launch::spmv, launch::jacobi_sweep, launch::invert_blocks, launch::multi_dot, launch::multi_axpy, launch::multi_dot1, launch::axpy_norm, launch::spmv_sweep0, launch::scale_from, launch::gmres_begin, launch::givens_step, launch::gmres_solve_y, launch::gmres_state_doubles, launch::sub,
        launch::axpy_state, launch::axpy_norm_givens, launch::limiter_michalak};
    return t;
}

}  // namespace AFX_NS
}  // namespace afx
