// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE ONLY: C entry points around the reference's OWN classes, compiled from
// /root/reference into oracle/_ref/libslam_ref.so by oracle/Makefile.ref.  Nothing here restates reference logic: it only
// constructs the reference objects and copies their outputs out.
#include "orb_extractor.h"              // /root/reference/src/orb_features
#include "DBoW2/FORB.h"                 // /root/reference/third_party/DBoW2
#include "DBoW2/TemplatedVocabulary.h"

#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>
#include <sys/mman.h>

// dropin/Makefile compiles this file too: the same entry points around the DROP-IN classes (dropin/orb_features/orb_extractor.h
// is then the "orb_extractor.h" found first), on the same heap, so that both libraries run the reference's Frame / KeyFrame /
// MapPoint / DBoW2 code in the same environment.
// ---- a monotonic heap for everything this library allocates --------------------------------------------------------------
// DistributeOctTree sorts pair<int, ExtractorNode*> (orb_extractor.cpp:625), so nodes of equal size are ordered by their HEAP
// ADDRESS: with a general-purpose malloc that order depends on which freed chunks get reused.  This library is linked with
// -Bsymbolic-functions and replaces operator new by a bump allocator, so that addresses grow in allocation order and the
// tie-break becomes "later-created node = higher address" -- the definition the oracle and the CUDA path use (DESIGN.md
// section 2, item 1).  The C++ runtime is linked statically into this library (Makefile.ref), so every allocation of the reference code and of the
// library internals it calls goes through these operators; memory is only reclaimed by the entry points that rewind.
namespace {
char* g_base = nullptr;
std::atomic<size_t> g_off(0);
#ifndef REF_ARENA_BITS
#define REF_ARENA_BITS 36
#endif
const size_t kArena = (size_t)1 << REF_ARENA_BITS;  // virtual reservation, committed lazily (dropin/Makefile asks for less: both
                                                    // libraries live in one test process)
std::once_flag g_once;
std::atomic<int> g_malloc_mode(0);  // > 0: plain malloc/free (throughput runs: the arena never reclaims memory)
inline void* arena_alloc(size_t n) {  // thread-safe: Frame's stereo constructor extracts on two std::threads (frame.cpp:86-89)
  std::call_once(g_once, []() {
    void* p = mmap(nullptr, kArena, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    g_base = p == MAP_FAILED ? nullptr : static_cast<char*>(p);
  });
  if (!g_base) return nullptr;
  const size_t need = (n + 15) & ~(size_t)15;
  const size_t a = g_off.fetch_add(need);
  if (a + need > kArena) return nullptr;
  return g_base + a;
}
inline bool in_arena(void* p) { return g_base && p >= g_base && p < g_base + kArena; }
}  // namespace
// rewinds the arena on scope exit: only for entry points whose allocations are all dead by then (ref_orb_extract)
struct ArenaScope {
  size_t mark;
  ArenaScope() : mark(g_off.load()) {}
  ~ArenaScope() { g_off.store(mark); }
};
static inline void* any_alloc(size_t n) {
  void* p = g_malloc_mode.load(std::memory_order_relaxed) > 0 ? malloc(n ? n : 1) : arena_alloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
// the replacements serve this library only: oracle/hide_new.map keeps them out of the dynamic symbol table
#define REF_HIDDEN
REF_HIDDEN void* operator new(size_t n) { return any_alloc(n); }
REF_HIDDEN void* operator new[](size_t n) { return any_alloc(n); }
extern "C" void ref_set_malloc_mode(int on) { g_malloc_mode.store(on); }
REF_HIDDEN void operator delete(void* p) noexcept { if (p && !in_arena(p)) free(p); }
REF_HIDDEN void operator delete[](void* p) noexcept { if (p && !in_arena(p)) free(p); }
REF_HIDDEN void operator delete(void* p, size_t) noexcept { if (p && !in_arena(p)) free(p); }
REF_HIDDEN void operator delete[](void* p, size_t) noexcept { if (p && !in_arena(p)) free(p); }

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVocabulary;

extern "C" {

// ORBextractor::Compute (orb_extractor.cpp:985-1049) on a u8 image; returns the keypoint count (negative: capacity too small)
int ref_orb_extract(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, const unsigned char* img, int w, int h,
                    int stride, void* kps_out /* 28-byte cv::KeyPoint */, unsigned char* desc_out, int cap) {
  ArenaScope scope;
  ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
  cv::Mat image(h, w, CV_8UC1, const_cast<unsigned char*>(img), (size_t)stride);
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  ex.Compute(image, cv::Mat(), kps, desc);
  const int n = (int)kps.size();
  if (n > cap) return -n;
  if (n) {
    std::memcpy(kps_out, kps.data(), (size_t)n * sizeof(cv::KeyPoint));
    for (int i = 0; i < n; ++i) std::memcpy(desc_out + (size_t)i * 32, desc.ptr(i), 32);
  }
  return n;
}

// the extractor's scale tables (orb_extractor.cpp:356-390)
void ref_orb_tables(int nfeatures, float scale_factor, int nlevels, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2) {
  ORBextractor ex(nfeatures, scale_factor, nlevels, 20, 7);
  for (int i = 0; i < nlevels; ++i) {
    scale[i] = ex.GetScaleFactors()[i]; inv_scale[i] = ex.GetInverseScaleFactors()[i];
    sigma2[i] = ex.GetScaleSigmaSquares()[i]; inv_sigma2[i] = ex.GetInverseScaleSigmaSquares()[i];
  }
}

// OrbVocabulary::loadFromTextFile + transform(features, BowVector, FeatureVector, levelsup)
// (TemplatedVocabulary.h:1335-1422, 1124-1190); outputs flattened like orc_bow_transform
void* ref_voc_load_text(const char* path) {
  RefVocabulary* v = new RefVocabulary();
  if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
  return v;
}
void ref_voc_destroy(void* v) { delete static_cast<RefVocabulary*>(v); }
int ref_voc_transform(void* vp, int n, const unsigned char* desc, int levelsup, unsigned* bow_words, double* bow_values, int* n_bow,
                      unsigned* fv_nodes, int* fv_start, unsigned* fv_idx, int* n_fv) {
  RefVocabulary* V = static_cast<RefVocabulary*>(vp);
  std::vector<cv::Mat> feats;
  for (int i = 0; i < n; ++i) {
    cv::Mat d(1, 32, CV_8U);
    std::memcpy(d.data, desc + (size_t)i * 32, 32);
    feats.push_back(d);
  }
  DBoW2::BowVector bv;
  DBoW2::FeatureVector fv;
  V->transform(feats, bv, fv, levelsup);
  int u = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++u) { bow_words[u] = it->first; bow_values[u] = it->second; }
  *n_bow = u;
  int f = 0, p = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++f) {
    fv_nodes[f] = it->first;
    fv_start[f] = p;
    for (size_t k = 0; k < it->second.size(); ++k) fv_idx[p++] = it->second[k];
  }
  fv_start[f] = p;
  *n_fv = f;
  return u;
}

}  // extern "C"
