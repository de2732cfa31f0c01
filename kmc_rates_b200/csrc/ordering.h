// Level-structure sweeps of the nested-dissection ordering on the GPU (ordering.cu): the engine hands one of these to
// ngt::analyse_sparse, which uses it for the few large subgraphs at the top of the dissection tree, where the ordering
// has no tree parallelism yet and one serial sweep over 2*10^6 adjacency entries costs ~5 ms of host time.
#pragma once
#include <cstdint>

#include "symbolic.h"

namespace ngt {

// Process-wide accelerator for `device` (created on first use, kept until release_gpu_sweeps()); nullptr when the
// environment variable NGT_B200_NO_GPU_SWEEPS is set.  Thread-safe: concurrent branches of the dissection take
// separate contexts (stream + buffers); a call that finds none free declines and the caller sweeps on the host.
SweepAccel *gpu_sweeps(int device);

// frees the device buffers of every context (the handles' ngt_trim path)
void release_gpu_sweeps();

// bytes copied by the sweeps since process start (added to the library's H2D / D2H counters)
long long gpu_sweeps_h2d_bytes();
long long gpu_sweeps_d2h_bytes();

}  // namespace ngt
