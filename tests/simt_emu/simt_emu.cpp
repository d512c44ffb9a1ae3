// TEST INFRASTRUCTURE ONLY — host-side SIMT emulator behind smrt_b200/csrc/simt.h (see that header).
// One OS thread per CUDA thread, one block at a time; barriers are pthread barriers.
#include "simt.h"

#include <algorithm>
#include <memory>

thread_local simt_dim3 threadIdx;
thread_local simt_dim3 blockIdx;
simt_dim3 blockDim;
simt_dim3 gridDim;

namespace simt {
BlockState* g_block = nullptr;
static std::vector<unsigned char> g_dyn_smem;
static std::mutex g_group_mutex;
struct GroupBarrier {
  pthread_barrier_t bar;
  std::vector<uint64_t> mailbox = std::vector<uint64_t>(32, 0);
};
static std::map<std::pair<int, unsigned>, GroupBarrier*> g_groups;  // (warp, mask)

unsigned char* dynamic_smem(size_t) { return g_dyn_smem.data(); }

static GroupBarrier* group_of(int warp, unsigned mask) {
  std::lock_guard<std::mutex> lk(g_group_mutex);
  auto key = std::make_pair(warp, mask);
  auto it = g_groups.find(key);
  if (it != g_groups.end()) return it->second;
  auto* g = new GroupBarrier();
  pthread_barrier_init(&g->bar, nullptr, __builtin_popcount(mask));
  g_groups[key] = g;
  return g;
}

void launch(unsigned grid, unsigned block, const std::function<void()>& body) {
  if (g_dyn_smem.size() < (1u << 20)) g_dyn_smem.assign(1u << 20, 0);
  gridDim.x = grid;
  blockDim.x = block;
  for (unsigned bid = 0; bid < grid; ++bid) {
    BlockState st;
    st.nthreads = (int)block;
    pthread_barrier_init(&st.block_barrier, nullptr, block);
    g_block = &st;
    for (auto& kv : g_groups) {
      pthread_barrier_destroy(&kv.second->bar);
      delete kv.second;
    }
    g_groups.clear();
    std::vector<std::thread> threads;
    threads.reserve(block);
    for (unsigned t = 0; t < block; ++t) {
      threads.emplace_back([&, t]() {
        threadIdx.x = t;
        blockIdx.x = bid;
        body();
      });
    }
    for (auto& th : threads) th.join();
    for (auto& kv : st.named) {
      pthread_barrier_destroy(kv.second);
      delete kv.second;
    }
    pthread_barrier_destroy(&st.block_barrier);
    g_block = nullptr;
  }
}
}  // namespace simt

void __syncthreads() { pthread_barrier_wait(&simt::g_block->block_barrier); }
void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

void smrt_named_barrier(int id, int nthreads) {
  pthread_barrier_t* bar;
  {
    std::lock_guard<std::mutex> lk(simt::g_block->named_mutex);
    auto key = std::make_pair(id, nthreads);
    auto it = simt::g_block->named.find(key);
    if (it == simt::g_block->named.end()) {
      bar = new pthread_barrier_t;
      pthread_barrier_init(bar, nullptr, nthreads);
      simt::g_block->named[key] = bar;
    } else {
      bar = it->second;
    }
  }
  pthread_barrier_wait(bar);
}

void simt_group_barrier(unsigned mask) {
  int lane = threadIdx.x & 31;
  if (!(mask & (1u << lane))) {
    std::fprintf(stderr, "simt: lane %d calls a warp primitive with mask %08x that excludes it\n", lane, mask);
    std::abort();
  }
  auto* g = simt::group_of(threadIdx.x >> 5, mask);
  pthread_barrier_wait(&g->bar);
}
void __syncwarp(unsigned mask) {
  if (__builtin_popcount(mask) > 1) simt_group_barrier(mask);
}

uint64_t simt_shfl_raw(unsigned mask, uint64_t v, int src_lane) {
  int lane = threadIdx.x & 31;
  if (!(mask & (1u << lane))) {
    std::fprintf(stderr, "simt: lane %d shuffles with mask %08x that excludes it\n", lane, mask);
    std::abort();
  }
  auto* g = simt::group_of(threadIdx.x >> 5, mask);
  g->mailbox[lane] = v;
  pthread_barrier_wait(&g->bar);
  uint64_t out = (mask & (1u << src_lane)) ? g->mailbox[src_lane] : v;
  pthread_barrier_wait(&g->bar);
  return out;
}

int __any_sync(unsigned mask, int pred) {
  int lane = threadIdx.x & 31;
  if (!(mask & (1u << lane))) {
    std::fprintf(stderr, "simt: lane %d votes with mask %08x that excludes it\n", lane, mask);
    std::abort();
  }
  auto* g = simt::group_of(threadIdx.x >> 5, mask);
  g->mailbox[lane] = pred ? 1 : 0;
  pthread_barrier_wait(&g->bar);
  int any = 0;
  int nlanes = simt::g_block->nthreads - (int)((threadIdx.x >> 5) << 5);
  for (int l = 0; l < 32 && l < nlanes; ++l)
    if (mask & (1u << l)) any |= (int)g->mailbox[l];
  pthread_barrier_wait(&g->bar);
  return any;
}

unsigned __ballot_sync(unsigned mask, int pred) {
  int lane = threadIdx.x & 31;
  auto* g = simt::group_of(threadIdx.x >> 5, mask);
  g->mailbox[lane] = pred ? 1 : 0;
  pthread_barrier_wait(&g->bar);
  unsigned out = 0;
  for (int l = 0; l < 32; ++l)
    if ((mask & (1u << l)) && g->mailbox[l]) out |= 1u << l;
  pthread_barrier_wait(&g->bar);
  return out;
}

unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  int lane = threadIdx.x & 31;
  auto* g = simt::group_of(threadIdx.x >> 5, mask);
  g->mailbox[lane] = v;
  pthread_barrier_wait(&g->bar);
  unsigned out = 0;
  for (int l = 0; l < 32; ++l)
    if (mask & (1u << l)) out = std::max(out, (unsigned)g->mailbox[l]);
  pthread_barrier_wait(&g->bar);
  return out;
}

int __syncthreads_or(int pred) {
  static std::atomic<int> flag{0};
  __syncthreads();
  if (threadIdx.x == 0) flag = 0;
  __syncthreads();
  if (pred) flag = 1;
  __syncthreads();
  return flag.load();
}
