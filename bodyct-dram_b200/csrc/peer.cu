// Small cross-GPU exchanges of the data-parallel training step over NVLink peer memory (new functionality,
// the reference has no distributed code).
//
// To be the single-process reference on the GLOBAL batch, every train-mode BatchNorm needs the batch statistics of all
// ranks: 14 (+2) layers x (forward: sum y, sum y^2, count; backward: sum dz, sum dz*xhat) = ~30 all-reduces of <= 1025
// doubles per step, each on the critical path between two kernels of one layer.  Through NCCL every one of them is a
// kernel launch of a general-purpose collective (ring / tree set-up, 15-25 us each on 8 GPUs) that also needs SMs the
// persistent convolution kernels occupy.  Here the exchange is ONE small kernel per call that does all of it:
//
//   every rank owns a MAILBOX in its own HBM, mapped into the address space of every peer (CUDA IPC over NVLink/NVSwitch):
//       data[slot][src_rank][kPeerMaxDoubles]   flag[slot][src_rank]   counter
//   call k (every rank issues the same sequence of calls; k is a device-side counter, so a replayed CUDA graph works):
//     1. push   : store the local payload into data[k % S][my_rank] of EVERY rank's mailbox (remote stores over NVLink)
//     2. signal : fence.sys, then st.release.sys  flag[k % S][my_rank] = k  in every rank's mailbox
//     3. wait   : ld.acquire.sys on the OWN mailbox until flag[k % S][q] >= k for every q (bounded spin, then trap)
//     4. reduce : out[i] = sum over q = 0..world-1 of data[k % S][q][i]  — fixed order, so all ranks get identical bits
//     5. (fused variant) the BatchNorm finalize of the layer: mean / rstd / scale / shift and the running statistics
//   A slot is reused S calls later; a rank can only be S calls ahead of a peer after that peer has finished reading the
//   slot (it needs the peer's flags of the S - 1 calls in between, which the peer writes from later kernels), so S >= 2 is
//   safe; S = 4.  One-shot all-gather + local reduce: latency = one NVLink store round (~2-3 us), no SM beyond one CTA.
//
// Testing without N GPUs: the kernel takes `nvirt` = number of ranks to play; with nvirt = world, block b plays rank b and
// all "peers" are buffers of one GPU (the blocks of ONE launch are co-resident and may wait on one another; separate
// kernels on one GPU may not, B200_PROFILING.md).
#include "common.cuh"

namespace dram {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerSlots = 4;
constexpr int kPeerMaxDoubles = 1040;       // 2 x 512 channels + count, rounded up

struct PeerMailbox {
  unsigned long long flag[kPeerSlots][kPeerMaxRanks];
  unsigned long long counter;
  unsigned long long pad[7];
  double data[kPeerSlots][kPeerMaxRanks][kPeerMaxDoubles];
};

struct PeerArgs {
  PeerMailbox* box[kPeerMaxRanks];           // box[q] = rank q's mailbox as mapped in THIS process
  const double* in[kPeerMaxRanks];           // per played rank (index = blockIdx.x)
  double* out[kPeerMaxRanks];
  int n, world, rank0;
};

struct BnFinalizeArgs {                      // per played rank; C == 0: plain all-reduce
  const float* gamma[kPeerMaxRanks];
  const float* beta[kPeerMaxRanks];
  float* running_mean[kPeerMaxRanks];
  float* running_var[kPeerMaxRanks];
  float* mean[kPeerMaxRanks];
  float* rstd[kPeerMaxRanks];
  float* scale[kPeerMaxRanks];
  float* shift[kPeerMaxRanks];
  double count[kPeerMaxRanks];               // local element count per channel: travels as payload element 2C
  float momentum, eps;
  int n_updates, C;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512)
k_peer_allreduce(const PeerArgs a, const BnFinalizeArgs f) {
  __shared__ unsigned long long seq_s;
  const int me = a.rank0 + blockIdx.x;
  PeerMailbox* mine = a.box[me];
  if (threadIdx.x == 0) seq_s = ++mine->counter;            // only this rank's kernels touch its counter, one at a time
  __syncthreads();
  const unsigned long long seq = seq_s;
  const int slot = (int)(seq % kPeerSlots);
  const double* in = a.in[blockIdx.x];
  // 1. push the payload into every rank's mailbox (own included)
  for (int q = 0; q < a.world; ++q) {
    double* dst = a.box[q]->data[slot][me];
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) dst[i] = (f.C > 0 && i == 2 * f.C) ? f.count[blockIdx.x] : in[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. signal, 3. wait
  if (threadIdx.x < a.world) {
    st_release_sys(&a.box[threadIdx.x]->flag[slot][me], seq);
    const unsigned long long* fl = &mine->flag[slot][threadIdx.x];
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (ld_acquire_sys(fl) < seq) {
      __nanosleep(64);
      if ((++spins & 0x3ff) == 0) {                         // wall-clock bound: a rank died or the ranks' call sequences diverged
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20ull * 1000000000ull) __trap();   // 20 s: fail the step instead of hanging the box
      }
    }
  }
  __syncthreads();
  // 4. reduce in rank order (identical bits on every rank); __ldcg: the lines were written by peers, never trust L1
  double* out = a.out[blockIdx.x];
  for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < a.world; ++q) s += __ldcg(&mine->data[slot][q][i]);
    out[i] = s;
  }
  if (f.C == 0) return;
  // 5. BatchNorm finalize on the GLOBAL sums (same arithmetic as k_bn_finalize); payload = [sum y (C), sum y^2 (C), count]
  __syncthreads();
  const int b = blockIdx.x, C = f.C;
  const double count = out[2 * C];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double m = out[c] / count;
    double var = out[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float r = (float)(1.0 / sqrt(var + (double)f.eps));
    const float g = f.gamma[b] ? f.gamma[b][c] : 1.f, be = f.beta[b] ? f.beta[b][c] : 0.f;
    f.mean[b][c] = (float)m;
    f.rstd[b][c] = r;
    f.scale[b][c] = g * r;
    f.shift[b][c] = be - (float)m * g * r;
    if (f.running_mean[b]) {
      const float unbiased = count > 1.0 ? (float)(var * count / (count - 1.0)) : (float)var;
      float rm = f.running_mean[b][c], rv = f.running_var[b][c];
      for (int i = 0; i < f.n_updates; ++i) {
        rm = (1.f - f.momentum) * rm + f.momentum * (float)m;
        rv = (1.f - f.momentum) * rv + f.momentum * unbiased;
      }
      f.running_mean[b][c] = rm;
      f.running_var[b][c] = rv;
    }
  }
}

}  // namespace dram

using namespace dram;

extern "C" {

size_t dram_peer_mailbox_bytes(void) { return sizeof(PeerMailbox); }
int dram_peer_max_doubles(void) { return kPeerMaxDoubles; }
int dram_peer_max_ranks(void) { return kPeerMaxRanks; }

int dram_peer_alloc(void** mailbox) {
  DRAM_REQUIRE(mailbox, "peer_alloc: bad arguments");
  DRAM_CUDA(cudaMalloc(mailbox, sizeof(PeerMailbox)));      // a plain cudaMalloc allocation: exportable through CUDA IPC
  DRAM_CUDA(cudaMemset(*mailbox, 0, sizeof(PeerMailbox)));
  DRAM_CUDA(cudaDeviceSynchronize());
  return DRAM_OK;
}

int dram_peer_free(void* mailbox) {
  if (mailbox) DRAM_CUDA(cudaFree(mailbox));
  return DRAM_OK;
}

int dram_peer_export(const void* mailbox, void* handle64) {
  DRAM_REQUIRE(mailbox && handle64, "peer_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  DRAM_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(mailbox)));
  return DRAM_OK;
}

int dram_peer_open(const void* handle64, void** mailbox) {
  DRAM_REQUIRE(handle64 && mailbox, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  DRAM_CUDA(cudaIpcOpenMemHandle(mailbox, h, cudaIpcMemLazyEnablePeerAccess));
  return DRAM_OK;
}

int dram_peer_close(void* mailbox) {
  if (mailbox) DRAM_CUDA(cudaIpcCloseMemHandle(mailbox));
  return DRAM_OK;
}

static int fill_args(PeerArgs& a, void* const* mailboxes, const double* const* in, double* const* out, int n, int rank, int world,
                     int nvirt) {
  DRAM_REQUIRE(mailboxes && in && out && n > 0 && n <= kPeerMaxDoubles, "peer_allreduce: payload of %d doubles (max %d)", n, kPeerMaxDoubles);
  DRAM_REQUIRE(world >= 1 && world <= kPeerMaxRanks && nvirt >= 1 && rank >= 0 && rank + nvirt <= world,
               "peer_allreduce: world %d (max %d), rank %d, %d played ranks", world, kPeerMaxRanks, rank, nvirt);
  for (int q = 0; q < kPeerMaxRanks; ++q) {
    a.box[q] = q < world ? reinterpret_cast<PeerMailbox*>(mailboxes[q]) : nullptr;
    a.in[q] = q < nvirt ? in[q] : nullptr;
    a.out[q] = q < nvirt ? out[q] : nullptr;
    DRAM_REQUIRE(q >= world || a.box[q], "peer_allreduce: mailbox of rank %d is NULL", q);
    DRAM_REQUIRE(q >= nvirt || (a.in[q] && a.out[q]), "peer_allreduce: in/out of played rank %d is NULL", q);
  }
  a.n = n; a.world = world; a.rank0 = rank;
  return DRAM_OK;
}

int dram_peer_allreduce_f64(void* const* mailboxes, const double* const* in, double* const* out, int n, int rank, int world,
                            int nvirt, void* stream) {
  PeerArgs a;
  int rc = fill_args(a, mailboxes, in, out, n, rank, world, nvirt);
  if (rc) return rc;
  BnFinalizeArgs f = {};
  k_peer_allreduce<<<nvirt, 512, 0, (cudaStream_t)stream>>>(a, f);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

int dram_bn_finalize_peer(void* const* mailboxes, const double* const* sums_in, const double* counts, double* const* sums_out,
                          int rank, int world, int nvirt, const float* const* gamma, const float* const* beta, float* const* running_mean,
                          float* const* running_var, float momentum, float eps, int n_updates, float* const* mean,
                          float* const* rstd, float* const* scale, float* const* shift, int C, void* stream) {
  DRAM_REQUIRE(C > 0 && 2 * C + 1 <= kPeerMaxDoubles, "bn_finalize_peer: C=%d does not fit the mailbox", C);
  DRAM_REQUIRE(counts && gamma && beta && running_mean && running_var && mean && rstd && scale && shift, "bn_finalize_peer: bad arguments");
  PeerArgs a;
  int rc = fill_args(a, mailboxes, sums_in, sums_out, 2 * C + 1, rank, world, nvirt);
  if (rc) return rc;
  BnFinalizeArgs f = {};
  for (int q = 0; q < nvirt; ++q) {
    f.gamma[q] = gamma[q]; f.beta[q] = beta[q]; f.running_mean[q] = running_mean[q]; f.running_var[q] = running_var[q];
    f.mean[q] = mean[q]; f.rstd[q] = rstd[q]; f.scale[q] = scale[q]; f.shift[q] = shift[q]; f.count[q] = counts[q];
    DRAM_REQUIRE(f.mean[q] && f.rstd[q] && f.scale[q] && f.shift[q], "bn_finalize_peer: outputs of played rank %d are NULL", q);
    DRAM_REQUIRE((f.running_mean[q] == nullptr) == (f.running_var[q] == nullptr), "bn_finalize_peer: running stats must both be set");
  }
  f.momentum = momentum; f.eps = eps; f.n_updates = n_updates; f.C = C;
  k_peer_allreduce<<<nvirt, 512, 0, (cudaStream_t)stream>>>(a, f);
  DRAM_LAUNCH_CHECK();
  return DRAM_OK;
}

}  // extern "C"
