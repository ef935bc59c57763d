// One-box multi-GPU stable radix sort behind the C ABI (b200rs_multi_comm_* / b200rs_sort_multi, include/b200rs.h).
//
// Reference entry point being replaced: cudax::sort over a communicator
// (/root/reference/cudax/include/cuda/experimental/__multi_gpu/algorithm/sort/sort.h:96-127, protocol in
// hss/execute.h:56-129, hss/histogramming.h:522-610, hss/data_exchange.h:380-450); C-ABI model:
// /root/reference/c/parallel/include/cccl/c/radix_sort.h:110-124 (two-phase temp storage, plain pointers).
//
// One process per GPU.  Every rank owns ONE device allocation [control block | key receive buffer | value receive
// buffer] that is mapped into every peer with CUDA IPC at communicator creation (the only step that needs the caller's
// out-of-band all-gather).  After that a sort is a fixed sequence of kernels on the caller's stream with NO host wait
// and NO NCCL call:
//
//   round 0 .. key_bytes-1 of the exact MSD radix select over the UNSORTED shard:
//       local histogram kernel (upsweep / select_histogram_kernel)
//       multi_round_kernel   -- the "all-reduce": every rank PUSHES its counters into its slot of every peer's
//                               control block over NVLink, releases a flag, waits for the peers' flags, sums the
//                               slots and picks the bin of every splitter on the device
//   (last round, same kernel) second push of (below, equal, prefix) -> every rank derives the exact cut points of every
//                               source rank, the bucket offsets and the per-destination receive addresses: the
//                               PartitionPlan, in device memory
//   fused partition + exchange pass (onesweep in bucket mode reading the plan; stores go straight into the peers'
//                               receive buffers)
//   multi_barrier_kernel      -- every source's stores have landed
//   b200rs_sort               -- ONE local stable sort of the received items into the caller's output
//
// Flags carry a sequence number that both sides advance in lockstep (every rank runs the same phases for the same key
// width); slots are double-buffered by sequence parity, so a fast rank can never overwrite counters a slow rank is
// still reading.  A wait that sees no signal for ~10 s gives up and records MULTI_ERR_TIMEOUT instead of hanging the
// GPU; b200rs_multi_status reports it.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/b200rs.h"
#include "common.cuh"

namespace b200rs
{
int partition_with_plan(void* d_temp_storage, size_t* temp_storage_bytes, const void* d_keys_in, const void* d_values_in,
                        uint64_t num_items, int key_kind, int key_bytes, int value_bytes, int descending, int num_dests,
                        const PartitionPlan* d_plan, cudaStream_t stream);

constexpr int MAX_RANKS = 16;
constexpr int MAX_T     = 15; // splitters = ranks - 1
using ull               = unsigned long long;

enum MultiErr : uint32_t
{
  MULTI_ERR_TIMEOUT      = 1, // a peer never signalled
  MULTI_ERR_CAPACITY     = 2, // some shard does not fit the receive buffers
  MULTI_ERR_INCONSISTENT = 4  // the select rounds do not add up (cannot happen unless memory was corrupted)
};

struct MultiSlot
{
  ull hist[MAX_T * RADIX];
  ull n_local;
  ull lt[MAX_T], eq[MAX_T], prefix[MAX_T];
  ull pad[2];
};

struct MultiCtrl
{
  ull flag[MAX_RANKS]; // flag[src] = last sequence number rank `src` released to this rank
  ull pad[16];
  MultiSlot slot[2][MAX_RANKS];
};

// Per-rank private state of one sort (device memory of the communicator)
struct MultiState
{
  ull target[MAX_T]; // global rank of each splitter: items that must end on lower ranks
  ull prefix[MAX_T], below[MAX_T], lt_local[MAX_T], eq_local[MAX_T];
  ull n_all[MAX_RANKS];
  uint32_t status; // MultiErr bits, sticky until b200rs_multi_status reads them
  uint32_t pad;
};

struct RoundArgs
{
  MultiCtrl* peers[MAX_RANKS]; // every rank's control block as mapped on THIS GPU (own block included)
  ull recv_keys[MAX_RANKS];    // every rank's receive buffers as mapped on this GPU
  ull recv_vals[MAX_RANKS];
  const ull* hist_local; // [rows][256]
  MultiState* state;
  PartitionPlan* plan;
  ull seq; // sequence number this kernel releases (the last round also uses seq + 1)
  ull n_local;
  ull capacity_bytes;
  int rank, world, rows, round, last;
  int es_k, es_v;
};

__device__ __forceinline__ void st_release_sys(ull* p, ull v)
{
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ ull ld_acquire_sys(const ull* p)
{
  ull v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ ull ld_sys(const ull* p) // data another GPU wrote into this GPU's memory: never from L1
{
  ull v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ ull global_ns()
{
  ull t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// release `seq` to every peer, then wait until every peer released at least `seq` to this rank
__device__ __forceinline__ void signal_and_wait(const RoundArgs& a, ull seq)
{
  // every thread's remote stores are ordered before the flag (fence by the storing threads, then the CTA barrier)
  __threadfence_system();
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < a.world)
  {
    st_release_sys(&a.peers[tid]->flag[a.rank], seq);
    const ull* mine = &a.peers[a.rank]->flag[tid];
    const ull t0    = global_ns();
    while (ld_acquire_sys(mine) < seq)
    {
      __nanosleep(200);
      if (global_ns() - t0 > 10000000000ull)
      {
        atomicOr(&a.state->status, uint32_t(MULTI_ERR_TIMEOUT));
        break;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ ull warp_excl_scan(ull v, ull& total)
{
  const uint32_t lane = threadIdx.x & 31;
  ull incl            = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1)
  {
    const ull n = __shfl_up_sync(0xffffffffu, incl, s);
    if (lane >= uint32_t(s))
    {
      incl += n;
    }
  }
  total = __shfl_sync(0xffffffffu, incl, 31);
  return incl - v;
}

// The exact cut points, bucket offsets and receive addresses from what every rank knows after the select rounds.
// lt[s][i] / eq[s][i]: items of source rank s strictly below / equal to splitter i; one thread.
__device__ void build_plan(const RoundArgs& a, int par)
{
  const int world = a.world, nt = world - 1, rank = a.rank;
  MultiState& st          = *a.state;
  PartitionPlan& plan     = *a.plan;
  const MultiCtrl* me     = a.peers[rank];
  uint32_t status         = 0;
  // edges[s][r]: first partitioned index of source s that goes to rank r (r = 0 .. world); only s <= rank is needed
  ull my_edges[MAX_RANKS + 1];
  ull dst_off[MAX_RANKS]; // items that lower source ranks send to rank r
  for (int r = 0; r < world; ++r)
  {
    dst_off[r] = 0;
  }
  for (int s = 0; s <= rank; ++s)
  {
    ull edges[MAX_RANKS + 1];
    edges[0]     = 0;
    edges[world] = st.n_all[s];
    for (int i = 0; i < nt; ++i)
    {
      ull lt_sum = 0, before = 0;
      for (int q = 0; q < world; ++q)
      {
        lt_sum += ld_sys(&me->slot[par][q].lt[i]);
        if (q < s)
        {
          before += ld_sys(&me->slot[par][q].eq[i]);
        }
      }
      const ull eq_s = ld_sys(&me->slot[par][s].eq[i]);
      const ull lt_s = ld_sys(&me->slot[par][s].lt[i]);
      // items EQUAL to the splitter that still go to lower ranks, handed out in source-rank order
      const ull need = st.target[i] > lt_sum ? st.target[i] - lt_sum : 0;
      ull take       = need > before ? need - before : 0;
      take           = take < eq_s ? take : eq_s;
      edges[i + 1]   = lt_s + take;
    }
    for (int r = 0; r < world; ++r)
    {
      if (edges[r + 1] < edges[r])
      {
        status |= MULTI_ERR_INCONSISTENT;
      }
      if (s < rank)
      {
        dst_off[r] += edges[r + 1] - edges[r];
      }
    }
    if (s == rank)
    {
      for (int r = 0; r <= world; ++r)
      {
        my_edges[r] = edges[r];
      }
    }
  }
  // distinct splitters, ascending (targets grow with i, so equal prefixes are adjacent); bucket sizes in bucket order:
  // (below u0), (== u0), (between u0 and u1), (== u1), ..., (above the last)
  int ns       = 0;
  ull prev_end = 0, running = 0;
  for (int b = 0; b < RADIX; ++b)
  {
    plan.bins[b] = 0;
  }
  for (int i = 0; i < nt; ++i)
  {
    if (i > 0 && st.prefix[i] == st.prefix[i - 1])
    {
      continue;
    }
    plan.splitters[ns] = st.prefix[i];
    const ull lt = st.lt_local[i], eq = st.eq_local[i];
    if (lt < prev_end)
    {
      status |= MULTI_ERR_INCONSISTENT;
    }
    plan.bins[2 * ns]     = running;
    running += lt - prev_end;
    plan.bins[2 * ns + 1] = running;
    running += eq;
    prev_end = lt + eq;
    ++ns;
  }
  plan.bins[2 * ns] = running;
  for (int i = ns; i < 16; ++i)
  {
    plan.splitters[i] = 0;
  }
  plan.num_splitters = uint32_t(ns);
  if (prev_end > a.n_local)
  {
    status |= MULTI_ERR_INCONSISTENT;
  }
  // receive addresses, biased so that partitioned index idx lands at ptr + idx * item size
  PeerTable& pt = plan.peer;
  pt.num_dests  = uint32_t(world);
  pt.pad        = 0;
  for (int r = 0; r < MAX_RANKS; ++r)
  {
    const ull bias      = r < world ? dst_off[r] - my_edges[r] : 0; // modulo 2^64
    pt.rank_dst_keys[r] = r < world ? a.recv_keys[r] + bias * ull(a.es_k) : 0;
    pt.rank_dst_vals[r] = (r < world && a.es_v > 0) ? a.recv_vals[r] + bias * ull(a.es_v) : 0;
    pt.seg_end[r]       = r + 1 < world ? uint32_t(my_edges[r + 1]) : 0xffffffffu;
  }
  const int nb = 2 * ns + 1;
  for (int b = 0; b < 32; ++b)
  {
    ull dk = 0, dv = 0;
    if (b < nb)
    {
      const ull lo = plan.bins[b];
      const ull hi = b + 1 < nb ? plan.bins[b + 1] : a.n_local;
      int r_lo = 0, r_hi = 0;
      for (int r = 0; r + 1 < world; ++r)
      {
        r_lo += my_edges[r + 1] <= lo ? 1 : 0;
        r_hi += (hi > lo && my_edges[r + 1] < hi) ? 1 : 0;
      }
      if (hi <= lo || r_lo == r_hi) // the whole bucket goes to one rank (else: resolved per item in the kernel)
      {
        dk = pt.rank_dst_keys[r_lo];
        dv = pt.rank_dst_vals[r_lo];
      }
    }
    pt.bucket_dst_keys[b] = dk;
    pt.bucket_dst_vals[b] = dv;
  }
  // every shard must fit the receive buffers (every rank sees the same n_all, so every rank decides alike)
  for (int r = 0; r < world; ++r)
  {
    const int es = a.es_k > a.es_v ? a.es_k : a.es_v;
    if (st.n_all[r] * ull(es) > a.capacity_bytes)
    {
      status |= MULTI_ERR_CAPACITY;
    }
  }
  plan.status = status; // non-zero: the partition pass exits without storing anything (onesweep_kernel, bucket mode)
  if (status != 0)
  {
    atomicOr(&st.status, status);
  }
}

// One select round: push the local counters, wait, reduce, pick.  One CTA.
__global__ void __launch_bounds__(1024) multi_round_kernel(const RoundArgs a)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int world = a.world, nt = world - 1;
  const int par       = int(a.seq & 1);
  const MultiCtrl* me = a.peers[a.rank];
  MultiState& st      = *a.state;

  // 1. push this rank's counters into its slot on every rank (own block included)
  const int total = a.rows * RADIX;
  for (int p = 0; p < world; ++p)
  {
    MultiSlot* dst = &a.peers[p]->slot[par][a.rank];
    for (int i = tid; i < total; i += blockDim.x)
    {
      dst->hist[i] = a.hist_local[i];
    }
    if (tid == 0 && a.round == 0)
    {
      dst->n_local = a.n_local;
    }
  }
  signal_and_wait(a, a.seq);

  // 2. round 0: every rank's item count -> the global rank of every splitter
  if (a.round == 0)
  {
    if (tid == 0)
    {
      ull run = 0;
      for (int r = 0; r < world; ++r)
      {
        const ull n = ld_sys(&me->slot[par][r].n_local);
        st.n_all[r] = n;
        run += n;
        if (r < nt)
        {
          st.target[r] = run;
        }
      }
    }
    if (tid < MAX_T)
    {
      st.prefix[tid] = st.below[tid] = st.lt_local[tid] = st.eq_local[tid] = 0;
    }
    __syncthreads();
  }

  // 3. one warp per splitter: first bin whose global running count reaches the (remaining) target rank
  if (warp < nt)
  {
    const int row = a.rows == 1 ? 0 : warp;
    ull g[8], l[8], gs = 0, ls = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      const int idx = row * RADIX + lane * 8 + j;
      l[j]          = a.hist_local[idx];
      ull s         = 0;
      for (int r = 0; r < world; ++r)
      {
        s += ld_sys(&me->slot[par][r].hist[idx]);
      }
      g[j] = s;
      gs += s;
      ls += l[j];
    }
    ull gtot, ltot;
    const ull gex  = warp_excl_scan(gs, gtot);
    const ull lex  = warp_excl_scan(ls, ltot);
    const ull tgt  = st.target[warp] > 0 ? st.target[warp] : 1;
    const ull want = tgt - st.below[warp];
    bool found     = false;
    ull cg = gex, cl = lex, g_before = 0, l_before = 0, eq = 0;
    uint32_t bin = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      if (!found && cg + g[j] >= want)
      {
        found    = true;
        bin      = uint32_t(lane * 8 + j);
        g_before = cg;
        l_before = cl;
        eq       = l[j];
      }
      cg += g[j];
      cl += l[j];
    }
    const uint32_t who = __ballot_sync(0xffffffffu, found);
    int src            = who != 0 ? __ffs(who) - 1 : 31;
    if (who == 0 && lane == 31) // target beyond the total (only when there are no items at all): last bin
    {
      bin      = RADIX - 1;
      g_before = cg - g[7];
      l_before = cl - l[7];
      eq       = l[7];
    }
    bin      = __shfl_sync(0xffffffffu, bin, src);
    g_before = __shfl_sync(0xffffffffu, g_before, src);
    l_before = __shfl_sync(0xffffffffu, l_before, src);
    eq       = __shfl_sync(0xffffffffu, eq, src);
    if (lane == 0)
    {
      st.below[warp] += g_before;
      st.lt_local[warp] += l_before;
      st.eq_local[warp] = eq;
      st.prefix[warp]   = st.prefix[warp] * RADIX + bin;
    }
  }
  if (!a.last)
  {
    return;
  }

  // 4. last round: all-gather (below, equal, prefix) the same way, then every rank builds its plan
  __syncthreads();
  const int par2 = int((a.seq + 1) & 1);
  if (tid < nt)
  {
    for (int p = 0; p < world; ++p)
    {
      MultiSlot* dst   = &a.peers[p]->slot[par2][a.rank];
      dst->lt[tid]     = st.lt_local[tid];
      dst->eq[tid]     = st.eq_local[tid];
      dst->prefix[tid] = st.prefix[tid];
    }
  }
  signal_and_wait(a, a.seq + 1);
  if (tid == 0)
  {
    build_plan(a, par2);
  }
}

// every rank's stores of the exchange have landed: release, then wait for every peer
__global__ void __launch_bounds__(32) multi_barrier_kernel(const RoundArgs a)
{
  signal_and_wait(a, a.seq);
}

} // namespace b200rs

using namespace b200rs;

struct b200rs_multi_comm
{
  int rank = 0, world = 1, device = 0;
  size_t capacity      = 0; // bytes of each receive buffer
  unsigned char* base  = nullptr;
  unsigned char* peer_base[MAX_RANKS] = {};
  MultiState* state    = nullptr;
  PartitionPlan* plan  = nullptr;
  ull* hist            = nullptr; // [MAX_T][256]
  ull seq              = 0;
  size_t off_keys = 0, off_vals = 0;
  int last_launches = 0;
  bool timing       = false; // record an event at every phase boundary of the next sorts
  cudaEvent_t ev[5] = {};
  int ev_used       = 0;
};

static void mark_phase(b200rs_multi_comm* c, int i, cudaStream_t stream)
{
  if (c->timing)
  {
    if (c->ev[i] == nullptr)
    {
      cudaEventCreate(&c->ev[i]);
    }
    cudaEventRecord(c->ev[i], stream);
    c->ev_used = i + 1;
  }
}

static size_t round_up(size_t x, size_t a)
{
  return (x + a - 1) / a * a;
}

extern "C" {

int b200rs_multi_comm_create(b200rs_multi_comm** out, int rank, int world, size_t receive_bytes,
                             b200rs_allgather_fn allgather, void* allgather_ctx)
{
  if (out == nullptr || world < 1 || world > MAX_RANKS || rank < 0 || rank >= world || (world > 1 && allgather == nullptr))
  {
    return int(cudaErrorInvalidValue);
  }
  b200rs_multi_comm* c = new (std::nothrow) b200rs_multi_comm();
  if (c == nullptr)
  {
    return int(cudaErrorMemoryAllocation);
  }
  c->rank     = rank;
  c->world    = world;
  c->capacity = round_up(receive_bytes > 0 ? receive_bytes : 1, 512);
  cudaError_t e = cudaGetDevice(&c->device);
  c->off_keys   = round_up(sizeof(MultiCtrl), 512);
  c->off_vals   = c->off_keys + c->capacity;
  if (e == cudaSuccess)
  {
    e = cudaMalloc(&c->base, c->off_vals + c->capacity);
  }
  if (e == cudaSuccess)
  {
    e = cudaMemset(c->base, 0, sizeof(MultiCtrl));
  }
  if (e == cudaSuccess)
  {
    e = cudaMalloc(&c->state, sizeof(MultiState));
  }
  if (e == cudaSuccess)
  {
    e = cudaMemset(c->state, 0, sizeof(MultiState));
  }
  if (e == cudaSuccess)
  {
    e = cudaMalloc(&c->plan, sizeof(PartitionPlan));
  }
  if (e == cudaSuccess)
  {
    e = cudaMalloc(&c->hist, sizeof(ull) * MAX_T * RADIX);
  }
  if (e == cudaSuccess)
  {
    e = cudaDeviceSynchronize(); // the control block is zero before any peer can signal into it
  }
  c->peer_base[rank] = c->base;
  if (e == cudaSuccess && world > 1)
  {
    // exchange the IPC handles through the caller's all-gather, map every peer's allocation
    struct Wire
    {
      cudaIpcMemHandle_t handle;
      unsigned long long capacity;
      int ok;
      int pad;
    };
    Wire mine;
    memset(&mine, 0, sizeof(mine));
    mine.capacity = c->capacity;
    mine.ok       = cudaIpcGetMemHandle(&mine.handle, c->base) == cudaSuccess ? 1 : 0;
    Wire all[MAX_RANKS];
    memset(all, 0, sizeof(all));
    if (allgather(allgather_ctx, &mine, all, sizeof(Wire)) != 0)
    {
      e = cudaErrorUnknown;
    }
    for (int r = 0; r < world && e == cudaSuccess; ++r)
    {
      if (!all[r].ok || all[r].capacity != c->capacity)
      {
        e = cudaErrorInvalidValue; // every rank must ask for the same receive size
      }
    }
    for (int r = 0; r < world && e == cudaSuccess; ++r)
    {
      if (r == rank)
      {
        continue;
      }
      void* p = nullptr;
      e       = cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess);
      c->peer_base[r] = static_cast<unsigned char*>(p);
    }
    // nobody signals before everybody has mapped everybody (a second all-gather as the barrier)
    int token = e == cudaSuccess ? 1 : 0, tokens[MAX_RANKS] = {};
    if (allgather(allgather_ctx, &token, tokens, sizeof(int)) != 0 && e == cudaSuccess)
    {
      e = cudaErrorUnknown;
    }
    for (int r = 0; r < world && e == cudaSuccess; ++r)
    {
      if (!tokens[r])
      {
        e = cudaErrorUnknown;
      }
    }
  }
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    b200rs_multi_comm_destroy(c);
    return int(e);
  }
  *out = c;
  return 0;
}

int b200rs_multi_comm_destroy(b200rs_multi_comm* c)
{
  if (c == nullptr)
  {
    return 0;
  }
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
  {
    if (r != c->rank && c->peer_base[r] != nullptr)
    {
      cudaIpcCloseMemHandle(c->peer_base[r]);
    }
  }
  cudaFree(c->base);
  cudaFree(c->state);
  cudaFree(c->plan);
  cudaFree(c->hist);
  for (cudaEvent_t ev : c->ev)
  {
    if (ev != nullptr)
    {
      cudaEventDestroy(ev);
    }
  }
  cudaGetLastError();
  delete c;
  return 0;
}

int b200rs_multi_status(b200rs_multi_comm* c, int* status)
{
  if (c == nullptr || status == nullptr)
  {
    return int(cudaErrorInvalidValue);
  }
  uint32_t s    = 0;
  cudaError_t e = cudaMemcpy(&s, &c->state->status, sizeof(s), cudaMemcpyDeviceToHost); // waits for the device
  if (e == cudaSuccess && s != 0)
  {
    e = cudaMemset(&c->state->status, 0, sizeof(s));
  }
  *status = int(s);
  return int(e);
}

int b200rs_multi_last_launch_count(b200rs_multi_comm* c)
{
  return c != nullptr ? c->last_launches : 0;
}

int b200rs_multi_timing_enable(b200rs_multi_comm* c, int on)
{
  if (c == nullptr)
  {
    return int(cudaErrorInvalidValue);
  }
  c->timing  = on != 0;
  c->ev_used = 0;
  return 0;
}

int b200rs_multi_timing_read(b200rs_multi_comm* c, float* ms4)
{
  if (c == nullptr || ms4 == nullptr || c->ev_used < 5)
  {
    return int(cudaErrorInvalidValue);
  }
  cudaError_t e = cudaEventSynchronize(c->ev[4]);
  for (int i = 0; i < 4 && e == cudaSuccess; ++i)
  {
    e = cudaEventElapsedTime(&ms4[i], c->ev[i], c->ev[i + 1]);
  }
  return int(e);
}

int b200rs_sort_multi(
  b200rs_multi_comm* c,
  void* d_temp_storage,
  size_t* temp_storage_bytes,
  const void* d_keys_in,
  void* d_keys_out,
  const void* d_values_in,
  void* d_values_out,
  uint64_t num_items,
  int key_kind,
  int key_bytes,
  int value_bytes,
  int descending,
  b200rs_stream_t stream_)
{
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (c == nullptr || temp_storage_bytes == nullptr || key_kind < 0 || key_kind > 2)
  {
    return int(cudaErrorInvalidValue);
  }
  const int bits = key_bytes * 8;
  if (c->world == 1)
  {
    return b200rs_sort(d_temp_storage, temp_storage_bytes, d_keys_in, d_keys_out, d_values_in, d_values_out, num_items,
                       key_kind, key_bytes, value_bytes, 0, bits, descending, 0, nullptr, stream_);
  }
  // the fused partition pass exists for 4- and 8-byte keys with 0-, 4- or 8-byte values
  size_t part_bytes = 0;
  if (int rc = partition_with_plan(nullptr, &part_bytes, nullptr, nullptr, num_items, key_kind, key_bytes, value_bytes,
                                   descending, c->world, c->plan, stream))
  {
    return rc;
  }
  size_t sort_bytes = 0;
  if (int rc = b200rs_sort(nullptr, &sort_bytes, nullptr, nullptr, nullptr, nullptr, num_items, key_kind, key_bytes,
                           value_bytes, 0, bits, descending, 0, nullptr, stream_))
  {
    return rc;
  }
  // temp blob: [candidate state | candidates | max(partition temp, sort temp)]
  const uint64_t cand_cap   = num_items / 8 + (uint64_t(1) << 16);
  const size_t off_cstate   = 0;
  const size_t off_cand     = round_up(1026 * sizeof(ull), 256);
  const size_t off_work     = off_cand + round_up(size_t(cand_cap) * size_t(key_bytes), 256);
  const size_t total        = off_work + (part_bytes > sort_bytes ? part_bytes : sort_bytes) + 255;
  if (d_temp_storage == nullptr)
  {
    *temp_storage_bytes = total;
    return 0;
  }
  if (*temp_storage_bytes < total)
  {
    return int(cudaErrorInvalidValue);
  }
  if (num_items > 0
      && (d_keys_in == nullptr || d_keys_out == nullptr || (value_bytes > 0 && (d_values_in == nullptr || d_values_out == nullptr))))
  {
    return int(cudaErrorInvalidValue);
  }
  // A shard that does not fit the receive buffers: this rank still runs the select rounds and the barrier, so that no
  // peer waits for it in vain, skips its own data passes and reports the error; the peers see the same item counts,
  // flag MULTI_ERR_CAPACITY in build_plan and their partition passes exit without storing anything.
  const bool fits = num_items * uint64_t(key_bytes > value_bytes ? key_bytes : value_bytes) <= c->capacity;
  unsigned char* tb = reinterpret_cast<unsigned char*>(round_up(reinterpret_cast<size_t>(d_temp_storage), 256));
  ull* cstate       = reinterpret_cast<ull*>(tb + off_cstate);
  void* cand        = tb + off_cand;
  unsigned char* wk = tb + off_work;

  RoundArgs a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < c->world; ++r)
  {
    a.peers[r]     = reinterpret_cast<MultiCtrl*>(c->peer_base[r]);
    a.recv_keys[r] = ull(reinterpret_cast<uintptr_t>(c->peer_base[r] + c->off_keys));
    a.recv_vals[r] = ull(reinterpret_cast<uintptr_t>(c->peer_base[r] + c->off_vals));
  }
  a.hist_local     = c->hist;
  a.state          = c->state;
  a.plan           = c->plan;
  a.n_local        = num_items;
  a.capacity_bytes = c->capacity;
  a.rank           = c->rank;
  a.world          = c->world;
  a.es_k           = key_bytes;
  a.es_v           = value_bytes;
  const int nt     = c->world - 1;
  int launches     = 0;
  cudaError_t e    = cudaSuccess;

  mark_phase(c, 0, stream);
  for (int rnd = 0; rnd < key_bytes; ++rnd)
  {
    if (rnd == 0)
    {
      if (num_items > 0)
      {
        if (int rc = b200rs_digit_histogram(d_keys_in, num_items, key_kind, key_bytes, bits - 8, bits, descending,
                                            reinterpret_cast<uint64_t*>(c->hist), stream_))
        {
          return rc;
        }
        ++launches;
      }
      else if ((e = cudaMemsetAsync(c->hist, 0, sizeof(ull) * RADIX, stream)) != cudaSuccess)
      {
        return int(e);
      }
    }
    else
    {
      // the first full scan compacts the keys that can still matter; later rounds only look at those
      const bool emit = rnd == 1, use = rnd > 1;
      if (int rc = b200rs_select_histogram(
            d_keys_in, num_items, key_kind, key_bytes, descending, reinterpret_cast<const uint64_t*>(c->state->prefix), nt,
            rnd, reinterpret_cast<uint64_t*>(c->hist), use ? cand : nullptr,
            use ? reinterpret_cast<const uint64_t*>(cstate) : nullptr, emit ? cand : nullptr,
            emit ? reinterpret_cast<uint64_t*>(cstate) : nullptr, cand_cap, stream_))
      {
        return rc;
      }
      launches += num_items > 0 ? 1 : 0;
    }
    a.rows  = rnd == 0 ? 1 : nt;
    a.round = rnd;
    a.last  = rnd == key_bytes - 1 ? 1 : 0;
    a.seq   = ++c->seq;
    if (a.last)
    {
      ++c->seq; // the last round releases two sequence numbers
    }
    multi_round_kernel<<<1, 1024, 0, stream>>>(a);
    if ((e = cudaPeekAtLastError()) != cudaSuccess)
    {
      return int(e);
    }
    ++launches;
  }
  mark_phase(c, 1, stream);
  // fused partition + exchange: stores go straight into the peers' receive buffers (the previous sort's readers of
  // those buffers are done: every rank's round-0 release above is stream-ordered after its previous final sort)
  if (num_items > 0 && fits)
  {
    size_t pb = part_bytes;
    if (int rc = partition_with_plan(wk, &pb, d_keys_in, d_values_in, num_items, key_kind, key_bytes, value_bytes,
                                     descending, c->world, c->plan, stream))
    {
      return rc;
    }
    ++launches;
  }
  mark_phase(c, 2, stream);
  a.seq = ++c->seq;
  multi_barrier_kernel<<<1, 32, 0, stream>>>(a);
  if ((e = cudaPeekAtLastError()) != cudaSuccess)
  {
    return int(e);
  }
  ++launches;
  mark_phase(c, 3, stream);
  // ONE local stable sort of the received items (source-rank order + stable sort == global stable order)
  if (num_items > 0 && fits)
  {
    size_t sb = sort_bytes;
    if (int rc = b200rs_sort(wk, &sb, c->base + c->off_keys, d_keys_out, value_bytes > 0 ? c->base + c->off_vals : nullptr,
                             d_values_out, num_items, key_kind, key_bytes, value_bytes, 0, bits, descending, 0, nullptr,
                             stream_))
    {
      return rc;
    }
    const int n = b200rs_last_launch_count();
    launches += n > 1 ? n - 1 : n;
  }
  mark_phase(c, 4, stream);
  c->last_launches = launches;
  return fits ? 0 : int(cudaErrorMemoryAllocation);
}

} // extern "C"
