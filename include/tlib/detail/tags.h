// tlib/detail/tags.h -- policy tags of the tlib::ttv interface (B200-native drop-in).
//
// Same tag names as the reference (bassoy/ttv include/tlib/detail/tags.h:21-63) so that call sites such as
//     ttv(execution_policy::par_loop, slicing_policy::subtensor, fusion_policy::all, q, p, ...)
// keep compiling.  On the CPU these tags pick one of 19 loop-nest overloads; here every combination selects the same GPU
// path -- the tags are forwarded to the C-ABI as hints (`code` = enum ttv_b200_execution / _slicing / _fusion).
#pragma once

#include "../../ttv_b200.h"

namespace tlib::ttv {

namespace execution_policy {
struct sequential_t         { static constexpr int code = TTV_B200_EXEC_SEQ; };
struct sequential_blas_t    { static constexpr int code = TTV_B200_EXEC_SEQ_BLAS; };
struct parallel_t           { static constexpr int code = TTV_B200_EXEC_PAR; };
struct parallel_loop_t      { static constexpr int code = TTV_B200_EXEC_PAR_LOOP; };
struct parallel_taskloop_t  { static constexpr int code = TTV_B200_EXEC_PAR_TASKLOOP; };
struct parallel_task_t      { static constexpr int code = TTV_B200_EXEC_PAR_TASK; };
struct parallel_blas_t      { static constexpr int code = TTV_B200_EXEC_PAR_BLAS; };
struct parallel_loop_blas_t { static constexpr int code = TTV_B200_EXEC_PAR_BLAS_LOOP; };

inline constexpr sequential_t         seq{};
inline constexpr sequential_blas_t    seq_blas{};
inline constexpr parallel_t           par{};
inline constexpr parallel_loop_t      par_loop{};
inline constexpr parallel_taskloop_t  par_taskloop{};
inline constexpr parallel_task_t      par_task{};
inline constexpr parallel_blas_t      par_blas{};
inline constexpr parallel_loop_blas_t par_blas_loop{};
} // namespace execution_policy

namespace slicing_policy {
struct slice_t     { static constexpr int code = TTV_B200_SLICE; };
struct subtensor_t { static constexpr int code = TTV_B200_SUBTENSOR; };

inline constexpr slice_t     slice{};
inline constexpr subtensor_t subtensor{};
} // namespace slicing_policy

namespace fusion_policy {
struct none_t  { static constexpr int code = TTV_B200_FUSE_NONE; };
struct outer_t { static constexpr int code = TTV_B200_FUSE_OUTER; };
struct all_t   { static constexpr int code = TTV_B200_FUSE_ALL; };

inline constexpr none_t  none{};
inline constexpr outer_t outer{};
inline constexpr all_t   all{};
} // namespace fusion_policy

} // namespace tlib::ttv

// The names used by the reference's README and examples (README.md:64,79; example/interface1.cpp:42).
namespace tlib {
namespace execution {
inline constexpr ttv::execution_policy::sequential_t    seq{};
inline constexpr ttv::execution_policy::parallel_t      par{};
inline constexpr ttv::execution_policy::parallel_loop_t blas{};
}
namespace slicing {
inline constexpr ttv::slicing_policy::slice_t     small{};
inline constexpr ttv::slicing_policy::subtensor_t large{};
}
namespace loop_fusion {
inline constexpr ttv::fusion_policy::none_t  none{};
inline constexpr ttv::fusion_policy::outer_t outer{};
inline constexpr ttv::fusion_policy::all_t   all{};
}
} // namespace tlib
