// protocol_model.cc -- host-thread models of the three polled exchange protocols of the kernels
// (test infrastructure; nothing of the product links it).  compute-sanitizer's racecheck is not
// available on the GPU pool, so the LOGIC of the protocols -- is a slot ever overwritten before its
// last reader is done?  can a tag ever match stale data?  does every participant see the same sums
// in the same order? -- is exercised here with std::thread standing in for a CTA (or a rank), C++
// relaxed atomics standing in for ld/st.relaxed.{gpu,sys}, random scheduling jitter, and
// ThreadSanitizer watching the one place where plain (non-atomic) data is handed over.
//
//   model 1  AllGatherLL     nlo_kernels.cu PeerAllReduce / ReduceAndExchange / GatherLL: every
//                            participant stores its record as "LL" words (tag << 32 | 32 payload bits)
//                            into the slot [parity][source] of EVERY participant and polls its own
//                            slots until the tags match; slots double-buffered by the parity of the
//                            exchange number.  Also the cluster-partial all-gather of the resident
//                            kernel (one shared array instead of one per participant).
//   model 2  LeaderPublish   the streaming kernel's persistent grid: per-CTA partials (PLAIN stores,
//                            double-buffered by iteration parity) + an arrival counter; CTA 0 waits
//                            for all arrivals, sums in CTA order, steps, publishes the state as LL
//                            words in a SINGLE buffer; the other CTAs poll those words.
//
//   model 3  TileRing        the streaming kernel's TMA / mbarrier ring (gn_iteration_kernel): thread 0
//                            doubles as the producer and keeps STAGES - 1 bulk copies in flight, the
//                            warps consume a stage after its `full` barrier and release it on its
//                            `empty` barrier; ring positions are carried incrementally across
//                            iterations, the first tiles of the NEXT iteration are requested before
//                            the reduction, a loop that ends early waits for them before the CTA exits,
//                            and a share that fits the ring stays resident.  Software mbarriers
//                            (phase parity), a "DMA" thread that lands the copies late, plain stage
//                            memory: a stage overwritten before every warp released it, a tile read
//                            before it landed, a wrong ring position or a copy still in flight at exit
//                            fail the run (or show as a ThreadSanitizer report).
//
// A violated invariant shows as a wrong sum (a payload half of another iteration), a poll that
// never ends (the tag was overwritten: reported as a timeout) or a ThreadSanitizer report.
// Usage: protocol_model <participants> <iterations> <seed>; prints "PROTOCOL_MODEL_OK" on success.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace {

using Clock = std::chrono::steady_clock;
constexpr int kWords = 28;  // doubles per record (kAcc6)

uint64_t SplitMix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
// value `w` of participant `who` in exchange `seq` (any finite double will do)
double ValueOf(int who, uint64_t seq, int w) {
  const uint64_t h = SplitMix((static_cast<uint64_t>(who) << 40) ^ (seq << 8) ^ static_cast<uint64_t>(w));
  return static_cast<double>(static_cast<int64_t>(h >> 11)) * (1.0 / 9007199254740992.0) - 0.5;
}
uint64_t Bits(double v) {
  uint64_t b;
  memcpy(&b, &v, 8);
  return b;
}
double FromBits(uint64_t b) {
  double v;
  memcpy(&v, &b, 8);
  return v;
}
void Jitter(uint64_t* rng) {
  *rng = SplitMix(*rng);
  const unsigned r = static_cast<unsigned>(*rng & 63u);
  if (r == 0) std::this_thread::sleep_for(std::chrono::microseconds(50));  // a rank that falls behind
  else if (r < 8) std::this_thread::yield();
}
// PROTOCOL_MODEL_BREAK=1: the negative control -- model 1 without the parity double buffer (every exchange
// uses slot 0) and a short deadline; a participant one exchange ahead then overwrites words a slower one
// still waits for, and the run must FAIL.
bool g_break = false;
bool g_break_empty = false;  // PROTOCOL_MODEL_BREAK=3: model 3, the producer does not wait for `empty` -- must FAIL
bool g_break_ring = false;  // PROTOCOL_MODEL_BREAK=2: model 3 without the drain of the prefetch at loop exit -- must FAIL
struct Deadline {
  Clock::time_point end = Clock::now() + std::chrono::seconds((g_break || g_break_ring || g_break_empty) ? 2 : 60);
  bool Passed() const { return Clock::now() > end; }
};

std::atomic<int> g_failures{0};
void FailMsg(const char* what, int who, uint64_t seq) {
  if (g_failures.fetch_add(1) < 5) fprintf(stderr, "protocol_model: %s (participant %d, exchange %llu)\n", what, who,
                                           static_cast<unsigned long long>(seq));
}
// model 3: the other warps of a failed run sit in barriers that nobody will complete -- leave at once
[[noreturn]] void FailAndExit(const char* what, int a, uint64_t b) {
  fprintf(stderr, "protocol_model: %s (tiles %d, %llu)\n", what, a, static_cast<unsigned long long>(b));
  fflush(stderr);
  std::_Exit(1);
}

// ---------------------------------------------------------------- model 1
// slots[dest][parity][source][2 * kWords]
struct AllGatherLL {
  int n;
  std::vector<std::atomic<uint64_t>> words;
  explicit AllGatherLL(int participants)
      : n(participants), words(static_cast<size_t>(participants) * 2 * participants * 2 * kWords) {
    for (auto& w : words) w.store(0, std::memory_order_relaxed);
  }
  std::atomic<uint64_t>* Slot(int dest, int parity, int source) {
    return words.data() + ((static_cast<size_t>(dest) * 2 + parity) * n + source) * (2 * kWords);
  }
  // one participant's whole run: `iterations` exchanges numbered seq0 + 1 ...
  void Run(int me, uint64_t seq0, int iterations, uint64_t seed) {
    uint64_t rng = seed ^ (static_cast<uint64_t>(me) << 32);
    for (int it = 0; it < iterations; ++it) {
      const uint64_t seq = seq0 + static_cast<uint64_t>(it) + 1;
      const uint32_t tag = static_cast<uint32_t>(seq);
      const int parity = g_break ? 0 : static_cast<int>(seq & 1u);
      Jitter(&rng);  // the tile loop of this iteration
      for (int dest = 0; dest < n; ++dest) {
        std::atomic<uint64_t>* slot = Slot(dest, parity, me);
        for (int w = 0; w < kWords; ++w) {
          const uint64_t bits = Bits(ValueOf(me, seq, w));
          slot[2 * w].store((static_cast<uint64_t>(tag) << 32) | (bits & 0xffffffffULL), std::memory_order_relaxed);
          slot[2 * w + 1].store((static_cast<uint64_t>(tag) << 32) | (bits >> 32), std::memory_order_relaxed);
        }
        if ((rng >> 7) & 1u) Jitter(&rng);  // the stores to the peers do not land together
      }
      Deadline deadline;
      double total[kWords];
      for (int w = 0; w < kWords; ++w) total[w] = 0.0;
      for (int source = 0; source < n; ++source) {  // rank order: the same sum everywhere
        std::atomic<uint64_t>* slot = Slot(me, parity, source);
        for (int w = 0; w < kWords; ++w) {
          uint64_t lo, hi;
          unsigned polls = 0;
          while (true) {
            lo = slot[2 * w].load(std::memory_order_relaxed);
            hi = slot[2 * w + 1].load(std::memory_order_relaxed);
            if (static_cast<uint32_t>(lo >> 32) == tag && static_cast<uint32_t>(hi >> 32) == tag) break;
            if ((++polls & 0xfffu) == 0) {
              if (deadline.Passed()) {
                FailMsg("model 1: a tag never arrived (slot overwritten before it was read?)", me, seq);
                return;
              }
              std::this_thread::yield();
            }
          }
          total[w] += FromBits((hi << 32) | (lo & 0xffffffffULL));
        }
      }
      for (int w = 0; w < kWords; ++w) {
        double expect = 0.0;
        for (int source = 0; source < n; ++source) expect += ValueOf(source, seq, w);
        if (Bits(expect) != Bits(total[w])) {
          FailMsg("model 1: wrong sum (payload of another exchange)", me, seq);
          return;
        }
      }
    }
  }
};

// ---------------------------------------------------------------- model 2
struct LeaderPublish {
  int n;
  std::vector<double> partials;             // [2 parities][n][kWords], PLAIN memory (st.cg / ld.cg on the device)
  std::atomic<uint32_t> counter{0};         // arrivals, monotonic over the launch
  std::vector<std::atomic<uint64_t>> state; // [2 * kWords] LL words, tag = iteration + 1, SINGLE buffer
  explicit LeaderPublish(int ctas) : n(ctas), partials(static_cast<size_t>(2) * ctas * kWords, 0.0), state(2 * kWords) {
    for (auto& w : state) w.store(0, std::memory_order_relaxed);
  }
  // the "step": any deterministic function of the sums and of the iteration
  static double StepOf(double sum, int it, int w) { return sum * 0.5 + static_cast<double>(it) + 0.001 * w; }
  void Run(int me, int iterations, uint64_t seed) {
    uint64_t rng = seed ^ (static_cast<uint64_t>(me) << 32);
    double st[kWords];
    for (int w = 0; w < kWords; ++w) st[w] = 0.0;
    for (int it = 0; it < iterations; ++it) {
      const uint32_t want = static_cast<uint32_t>(it) + 1u;
      Jitter(&rng);  // tiles
      double* mine = partials.data() + (static_cast<size_t>(it & 1) * n + me) * kWords;
      for (int w = 0; w < kWords; ++w) mine[w] = ValueOf(me, static_cast<uint64_t>(it), w) + st[w] * 1e-3;
      // __syncthreads + __threadfence + atomicAdd: a releasing arrival
      counter.fetch_add(1u, std::memory_order_release);
      Deadline deadline;
      if (me == 0) {
        unsigned polls = 0;
        while (counter.load(std::memory_order_acquire) < want * static_cast<uint32_t>(n)) {
          if ((++polls & 0xfffu) == 0) {
            if (deadline.Passed()) { FailMsg("model 2: arrivals missing", me, it); return; }
            std::this_thread::yield();
          }
        }
        double next[kWords];
        for (int w = 0; w < kWords; ++w) {
          double s = 0.0;
          for (int c = 0; c < n; ++c) s += partials[(static_cast<size_t>(it & 1) * n + c) * kWords + w];  // CTA order
          next[w] = StepOf(s, it, w);
        }
        for (int w = 0; w < kWords; ++w) st[w] = next[w];
        for (int w = 0; w < kWords; ++w) {
          const uint64_t bits = Bits(st[w]);
          // (release / acquire on these words: on the device the order "leader has read the partials ->
          // leader publishes the state" is carried by the data dependency -- the state is computed from
          // the partials -- and "CTA has read the state -> CTA writes its next partial" by the loop's exit
          // condition and a bar.sync; C++ has no dependency ordering, release / acquire is its spelling)
          state[2 * w].store((static_cast<uint64_t>(want) << 32) | (bits & 0xffffffffULL), std::memory_order_release);
          state[2 * w + 1].store((static_cast<uint64_t>(want) << 32) | (bits >> 32), std::memory_order_release);
        }
      } else {
        for (int w = 0; w < kWords; ++w) {
          uint64_t lo, hi;
          unsigned polls = 0;
          while (true) {
            lo = state[2 * w].load(std::memory_order_acquire);
            hi = state[2 * w + 1].load(std::memory_order_acquire);
            if (static_cast<uint32_t>(lo >> 32) == want && static_cast<uint32_t>(hi >> 32) == want) break;
            if ((++polls & 0xfffu) == 0) {
              if (deadline.Passed()) { FailMsg("model 2: the state of an iteration never arrived", me, it); return; }
              std::this_thread::yield();
            }
          }
          st[w] = FromBits((hi << 32) | (lo & 0xffffffffULL));
        }
      }
      // every CTA must now hold the state the leader computed from ALL partials of this iteration;
      // the partials every CTA wrote are a function of (cta, it, previous state), so the check can be
      // replayed locally from the previous state of this CTA
      // (a stale or torn state propagates into the next partial and is caught one iteration later
      // by the checksum below at the latest)
      if ((rng & 0x300u) == 0) Jitter(&rng);
    }
    checksum[me] = 0.0;
    for (int w = 0; w < kWords; ++w) checksum[me] += st[w];
  }
  std::vector<double> checksum = std::vector<double>(1024, 0.0);
  // the same loop, sequentially
  static double Serial(int n, int iterations) {
    double st[kWords];
    for (int w = 0; w < kWords; ++w) st[w] = 0.0;
    for (int it = 0; it < iterations; ++it) {
      double next[kWords];
      for (int w = 0; w < kWords; ++w) {
        double s = 0.0;
        for (int c = 0; c < n; ++c) s += ValueOf(c, static_cast<uint64_t>(it), w) + st[w] * 1e-3;
        next[w] = StepOf(s, it, w);
      }
      for (int w = 0; w < kWords; ++w) st[w] = next[w];
    }
    double sum = 0.0;
    for (int w = 0; w < kWords; ++w) sum += st[w];
    return sum;
  }
};


// ---------------------------------------------------------------- model 3
struct SoftBarrier {  // mbarrier: arrival count + phase parity; wait(P) returns once the phase of parity P is over
  std::atomic<int> left{0};
  std::atomic<uint32_t> parity{0};
  int count = 1;
  void Init(int c) { count = c; left.store(c); parity.store(0); }
  void Arrive() {
    if (left.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      left.store(count, std::memory_order_relaxed);
      parity.fetch_xor(1u, std::memory_order_release);
    }
  }
  bool TryWait(uint32_t p) const { return parity.load(std::memory_order_acquire) != p; }
};

struct TileRing {
  static constexpr int kMaxStages = 4, kStageWords = 16;
  int warps, stages, my_tiles, iterations, exit_after;  // exit_after: the step of this iteration sets st.done (-1 never)
  bool solve_mode;
  SoftBarrier full[kMaxStages], empty[kMaxStages];
  long stage_data[kMaxStages][kStageWords];  // PLAIN memory: what a bulk copy writes and the warps read
  // the "TMA engine": copies land late and in order
  struct Copy { int stage; long content; };
  std::mutex mu;
  std::deque<Copy> queue;
  std::atomic<int> in_flight{0};
  std::atomic<bool> stop{false};
  std::atomic<int> done_flag{0};  // st.done, published behind the CTA barrier
  // a reusable CTA barrier (__syncthreads)
  std::atomic<int> bar_count{0};
  std::atomic<int> bar_gen{0};
  bool failed = false;
  int issued = 0, awaited = 0;  // bulk copies requested / whose landing warp 0 has waited for (warp 0 only)

  void CtaSync() {
    const int gen = bar_gen.load(std::memory_order_acquire);
    if (bar_count.fetch_add(1, std::memory_order_acq_rel) == warps - 1) {
      bar_count.store(0, std::memory_order_relaxed);
      bar_gen.fetch_add(1, std::memory_order_release);
    } else {
      while (bar_gen.load(std::memory_order_acquire) == gen) std::this_thread::yield();
    }
  }
  bool Wait(SoftBarrier& b, uint32_t parity, const char* what) {
    Deadline deadline;
    unsigned polls = 0;
    while (!b.TryWait(parity)) {
      if ((++polls & 0x3ffu) == 0) {
        if (deadline.Passed()) FailAndExit(what, my_tiles, static_cast<uint64_t>(stages));
        std::this_thread::yield();
      }
    }
    return true;
  }
  void Dma() {
    uint64_t rng = 99;
    while (true) {
      Copy c{-1, 0};
      {
        std::lock_guard<std::mutex> lock(mu);
        if (!queue.empty()) { c = queue.front(); queue.pop_front(); }
      }
      if (c.stage < 0) {
        if (stop.load()) return;
        std::this_thread::yield();
        continue;
      }
      Jitter(&rng);
      for (int k = 0; k < kStageWords; ++k) stage_data[c.stage][k] = c.content;
      in_flight.fetch_sub(1, std::memory_order_release);
      full[c.stage].Arrive();  // complete_tx: the phase of `full` ends when the bytes have landed
    }
  }
  // one warp of the CTA; warp 0 carries thread 0, the producer
  void Warp(int warp) {
    const bool producer = warp == 0;
    uint64_t rng = 7u + static_cast<uint64_t>(warp);
    int c_stage = 0, p_stage = 0;
    uint32_t c_phase = 0, p_phase = 0;
    const bool resident = iterations > 1 && my_tiles <= stages;
    int prefetched = 0;
    auto issue_tile = [&](int m, int for_iteration) {
      const int s = p_stage;
      const uint32_t phase = p_phase;
      if (++p_stage == stages) { p_stage = 0; p_phase ^= 1u; }
      if (!g_break_empty && !Wait(empty[s], phase ^ 1u, "model 3: producer stuck on an empty barrier")) return;
      in_flight.fetch_add(1, std::memory_order_relaxed);
      ++issued;
      std::lock_guard<std::mutex> lock(mu);
      queue.push_back(Copy{s, 1000L * for_iteration + m});
    };
    for (int it = 0; it < iterations; ++it) {
      if (done_flag.load(std::memory_order_acquire)) break;
      const bool need_load = !resident || it == 0;
      if (producer && need_load)
        for (int m = prefetched; m < stages - 1 && m < my_tiles; ++m) issue_tile(m, it);
      prefetched = 0;
      for (int m = 0; m < my_tiles; ++m) {
        const int s = resident ? m : c_stage;
        const uint32_t phase = c_phase;
        if (!resident && ++c_stage == stages) { c_stage = 0; c_phase ^= 1u; }
        if (need_load) {
          if (producer && m + stages - 1 < my_tiles) issue_tile(m + stages - 1, it);
          if (failed || !Wait(full[s], phase, "model 3: a warp stuck on a full barrier")) return;
          if (producer) ++awaited;
        }
        const long expect = 1000L * (resident ? 0 : it) + m;
        for (int k = 0; k < kStageWords; ++k)
          if (stage_data[s][k] != expect)
            FailAndExit("model 3: a warp read a stage that does not hold its tile", m, static_cast<uint64_t>(it));
        if (!resident) empty[s].Arrive();
        if ((rng & 3u) == 0) Jitter(&rng);
        rng = SplitMix(rng);
      }
      if (!resident && it + 1 < iterations && solve_mode) {
        const int ahead = my_tiles < stages - 1 ? my_tiles : stages - 1;
        if (producer)
          for (int m = 0; m < ahead; ++m) issue_tile(m, it + 1);
        prefetched = ahead;
      }
      CtaSync();  // reduction ...
      if (producer && it == exit_after) done_flag.store(1, std::memory_order_release);  // ... and the step
      CtaSync();
    }
    // a prefetch may still be in flight when the loop ends early: let it land before the CTA exits
    for (int m = 0; m < (g_break_ring ? 0 : prefetched); ++m) {
      if (!Wait(full[c_stage], c_phase, "model 3: the drain of the prefetch is stuck")) return;
      if (producer) ++awaited;
      if (++c_stage == stages) { c_stage = 0; c_phase ^= 1u; }
    }
  }
  bool Run() {
    for (int s = 0; s < kMaxStages; ++s) {
      full[s].Init(1);
      empty[s].Init(warps);
      for (int k = 0; k < kStageWords; ++k) stage_data[s][k] = -1;
    }
    std::thread dma([this]() { Dma(); });
    std::vector<std::thread> threads;
    for (int w = 0; w < warps; ++w) threads.emplace_back([this, w]() { Warp(w); });
    for (auto& t : threads) t.join();
    // the CTA has exited: nothing may still be on its way into its shared memory
    const int late = in_flight.load(std::memory_order_acquire);
    stop.store(true);
    dma.join();
    if ((late != 0 || issued != awaited) && !failed) {  // every copy requested has been waited for before the exit
      FailMsg("model 3: bulk copies not waited for when the CTA exits", my_tiles, static_cast<uint64_t>(issued - awaited));
      failed = true;
    }
    return !failed;
  }
};

}  // namespace

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 8;
  const int iterations = argc > 2 ? atoi(argv[2]) : 2000;
  const uint64_t seed = argc > 3 ? strtoull(argv[3], nullptr, 10) : 1;
  if (n < 1 || n > 1024 || iterations < 1) return 2;
  if (const char* b = getenv("PROTOCOL_MODEL_BREAK")) {
    g_break = b[0] == '1';
    g_break_ring = b[0] == '2';
    g_break_empty = b[0] == '3';
  }
  {
    // exchange numbers that cross the 32-bit wrap of the tag (the sequence number is 64-bit on the device,
    // the tag its low 32 bits)
    AllGatherLL model(n);
    const uint64_t seq0 = 0xffffffffULL - static_cast<uint64_t>(iterations / 2);
    std::vector<std::thread> threads;
    for (int r = 0; r < n; ++r) threads.emplace_back([&, r]() { model.Run(r, seq0, iterations, seed); });
    for (auto& t : threads) t.join();
  }
  {
    LeaderPublish model(n);
    std::vector<std::thread> threads;
    for (int c = 0; c < n; ++c) threads.emplace_back([&, c]() { model.Run(c, iterations, seed); });
    for (auto& t : threads) t.join();
    const double expect = LeaderPublish::Serial(n, iterations);
    for (int c = 0; c < n; ++c)
      if (Bits(model.checksum[static_cast<size_t>(c)]) != Bits(expect)) FailMsg("model 2: final state differs from the serial loop", c, 0);
  }
  if (!g_break) {
    int runs = 0;
    for (int stages = 2; stages <= TileRing::kMaxStages; ++stages)
      for (int my_tiles : {0, 1, 2, 3, 4, 5, 9, 23})
        for (int iters : {1, 2, 5})
          for (int exit_after : {-1, 0, 2})
            for (int solve_mode = 0; solve_mode < 2; ++solve_mode) {
              if (exit_after >= iters) continue;
              TileRing ring;
              ring.warps = 4;
              ring.stages = stages;
              ring.my_tiles = my_tiles;
              ring.iterations = iters;
              ring.exit_after = exit_after;
              ring.solve_mode = solve_mode != 0;
              ring.Run();
              ++runs;
              if (g_failures.load() != 0) goto ring_done;  // a failed run may leave waits that only end at their deadline
            }
  ring_done:
    if (g_failures.load() == 0) printf("TILE_RING_OK runs=%d\n", runs);
  }
  if (g_failures.load() != 0) {
    fprintf(stderr, "protocol_model: %d failure(s)\n", g_failures.load());
    return 1;
  }
  printf("PROTOCOL_MODEL_OK participants=%d iterations=%d\n", n, iterations);
  return 0;
}
