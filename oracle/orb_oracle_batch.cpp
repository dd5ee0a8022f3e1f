// orb_oracle_batch.cpp -- multi-threaded driver over the CPU oracle, used ONLY by bench.py's
// cpu_baseline / --impl reference legs (test infrastructure; see orb_oracle.h).
// Mirrors the reference's threading: 2 threads per stereo pair for extraction
// (src/data/frame.cpp:86-89), and one worker per host core over independent pairs for batches
// (the offline loop of examples/main_stereo.cpp:102-143 has no cross-frame state on this path).
#include "orb_oracle.h"

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

extern "C" {

// Processes n_pairs stereo pairs (left[i], right[i], each w*h tightly packed) with n_threads
// workers; returns wall seconds.  total_kps/total_matches accumulate checksums.
double orc_bench_stereo_batch(const uint8_t* const* left, const uint8_t* const* right, int n_pairs, int w, int h,
                              int nfeatures, float scale, int nlevels, int ini_th, int min_th, float bf,
                              float baseline, int n_threads, int pair_threads, long* total_kps,
                              long* total_matches) {
  std::atomic<int> next(0);
  std::atomic<long> kps_sum(0), match_sum(0);
  auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {
    orc_extractor* L = orc_extractor_create(nfeatures, scale, nlevels, ini_th, min_th);
    orc_extractor* R = orc_extractor_create(nfeatures, scale, nlevels, ini_th, min_th);
    const int cap = nfeatures + 64 * nlevels;
    std::vector<orc_keypoint> kl(cap), kr(cap);
    std::vector<uint8_t> dl((size_t)cap * 32), dr((size_t)cap * 32);
    std::vector<float> ur(cap), dp(cap);
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_pairs) break;
      int nl = 0, nr = 0;
      if (pair_threads >= 2) {
        std::thread tl([&]() { nl = orc_extract(L, left[i], w, h, w, kl.data(), dl.data(), cap); });
        std::thread tr([&]() { nr = orc_extract(R, right[i], w, h, w, kr.data(), dr.data(), cap); });
        tl.join();
        tr.join();
      } else {
        nl = orc_extract(L, left[i], w, h, w, kl.data(), dl.data(), cap);
        nr = orc_extract(R, right[i], w, h, w, kr.data(), dr.data(), cap);
      }
      if (nl < 0) nl = 0;
      if (nr < 0) nr = 0;
      const int m = orc_stereo_match(L, R, nl, kl.data(), dl.data(), nr, kr.data(), dr.data(), bf, baseline,
                                     ur.data(), dp.data());
      kps_sum += nl + nr;
      match_sum += m;
    }
    orc_extractor_destroy(L);
    orc_extractor_destroy(R);
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (total_kps) *total_kps = kps_sum.load();
  if (total_matches) *total_matches = match_sum.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
