// Gradient all-reduce issued by the library itself: an NCCL communicator owned by libpamnet, created
// from a unique id the host side broadcasts with torch.distributed.  model_backward then enqueues one ncclAllReduce
// (average) per gradient bucket on a communication stream the moment the bucket's weight gradients have been issued --
// no Python between the buckets (seven dist.all_reduce calls from Python cost more host time than the overlap saved:
// 2.05 vs 1.77 ms/step on 2 GPUs), only the last bucket is exposed, and the caller's stream waits for it before
// model_backward returns control of the gradients.  NCCL is resolved with dlopen at run time (the copy torch already
// loaded), so the library has no link-time dependency and still loads on a CPU-only box.
#pragma once
#include "common.cuh"

namespace pamnet {

int comm_unique_id(void* out128);
int comm_init(const void* id128, int rank, int world);
int comm_enable(int on);              // temporarily bypass the collective (single-rank profiling passes)
int comm_destroy();
bool comm_active();                   // a communicator exists for the current device and is enabled
cudaStream_t comm_stream();           // its stream
// average buf[0, count) over the ranks, in place, on the communication stream (the caller orders the stream)
int comm_allreduce_avg(float* buf, int64_t count);

}  // namespace pamnet
