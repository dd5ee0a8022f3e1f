// orbfe_bow.cu -- host side of the vocabulary / BoW entry points of include/orbfe.h (kernels in k_bow.cuh).
// N3 of SURVEY 8f: Frame::ComputeBoW (frame.cpp:258-263), KeyFrame::ComputeBoW (keyframe.cpp:127-137) ->
// DBoW2 TemplatedVocabulary::transform (third_party/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1250).
#include "../../include/orbfe.h"

#include "k_bow.cuh"
#include "orbfe_host.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <new>
#include <sstream>
#include <string>
#include <vector>

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return orbfe_fail(ORBFE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#ifdef ORBFE_EMU
#define BOW_LAUNCH(v, kernel, grid, block, ...) emu::launch(grid, block, 0, [&]() { kernel(__VA_ARGS__); })
#else
#define BOW_LAUNCH(v, kernel, grid, block, ...) kernel<<<grid, block, 0, (v)->stream>>>(__VA_ARGS__)
#endif

struct orbfe_vocabulary {
  int device = 0;
  cudaStream_t stream = nullptr;
  int k = 0, L = 0, scoring = 0, weighting = 0;
  int nNodes = 0, nWords = 0;
  uint8_t* d_desc = nullptr;
  int* d_childStart = nullptr;
  int* d_child = nullptr;
  double* d_weight = nullptr;
  unsigned* d_wordId = nullptr;
  // per-call scratch, grown on demand
  int cap = 0;  // features (power of two)
  uint8_t* d_feat = nullptr;
  unsigned *d_word = nullptr, *d_node = nullptr, *d_bowWord = nullptr, *d_fvNode = nullptr, *d_fvIdx = nullptr;
  double *d_fw = nullptr, *d_bowValue = nullptr;
  unsigned long long *d_wkey = nullptr, *d_nkey = nullptr;
  int *d_bowStart = nullptr, *d_fvStart = nullptr, *d_counts = nullptr;   // d_fvStart, d_counts and the arrays above: views into d_out
  // one transform = one H2D copy (descriptors from a pinned staging copy) and ONE D2H copy: every output array lives in one
  // device block (laid out per call for n features) with a pinned mirror
  char* d_out = nullptr;
  char* h_out = nullptr;    // pinned
  uint8_t* h_feat = nullptr;  // pinned
  VocabTree tree() const {
    VocabTree V;
    V.desc = d_desc; V.childStart = d_childStart; V.child = d_child; V.weight = d_weight; V.wordId = d_wordId;
    V.nNodes = nNodes; V.L = L;
    return V;
  }
};

template <class T>
static cudaError_t regrow(T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  return cudaMalloc(p, (count ? count : 1) * sizeof(T));
}

static int ensure_scratch(orbfe_vocabulary* v, int n) {
  int n2 = 1024;
  while (n2 < n) n2 <<= 1;
  if (n2 <= v->cap) return ORBFE_OK;
  CUDA_TRY(cudaStreamSynchronize(v->stream));
  const size_t c = (size_t)n2;
  CUDA_TRY(regrow(&v->d_feat, c * 32));
  CUDA_TRY(regrow(&v->d_fw, c));
  CUDA_TRY(regrow(&v->d_wkey, c)); CUDA_TRY(regrow(&v->d_nkey, c));
  CUDA_TRY(regrow(&v->d_bowStart, c + 1));
  CUDA_TRY(regrow(&v->d_out, 36 * c + 256));
  if (v->h_out) cudaFreeHost(v->h_out);
  if (v->h_feat) cudaFreeHost(v->h_feat);
  v->h_out = nullptr; v->h_feat = nullptr;
  CUDA_TRY(cudaMallocHost(&v->h_out, 36 * c + 256));
  CUDA_TRY(cudaMallocHost(&v->h_feat, c * 32));
  v->cap = n2;
  return ORBFE_OK;
}

extern "C" {

int orbfe_vocabulary_destroy(orbfe_vocabulary* v) {
  if (!v) return ORBFE_OK;
  cudaSetDevice(v->device);
  if (v->stream) cudaStreamSynchronize(v->stream);
  cudaFree(v->d_desc); cudaFree(v->d_childStart); cudaFree(v->d_child); cudaFree(v->d_weight); cudaFree(v->d_wordId);
  cudaFree(v->d_feat); cudaFree(v->d_fw); cudaFree(v->d_wkey); cudaFree(v->d_nkey);
  cudaFree(v->d_bowStart); cudaFree(v->d_out);
  if (v->h_out) cudaFreeHost(v->h_out);
  if (v->h_feat) cudaFreeHost(v->h_feat);
  if (v->stream) cudaStreamDestroy(v->stream);
  delete v;
  return ORBFE_OK;
}

// the tree exactly as TemplatedVocabulary::loadFromTextFile builds it (TemplatedVocabulary.h:1372-1417): node 0 is the
// root; node i >= 1 hangs under parent[i] (children in push_back order = ascending id); a node flagged as leaf takes the
// next word id
int orbfe_vocabulary_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                            const uint8_t* is_leaf, const uint8_t* desc, const double* weight, orbfe_vocabulary** out) {
  if (!out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (k < 0 || k > 20 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3)  // :1356
    return orbfe_fail(ORBFE_ERR_INVALID, "not a vocabulary header: k=%d L=%d scoring=%d weighting=%d", k, L, scoring, weighting);
  if (n_nodes < 1 || (n_nodes > 1 && (!parent || !is_leaf || !desc || !weight))) return orbfe_fail(ORBFE_ERR_INVALID, "bad node arrays");
  std::vector<int> cnt(n_nodes + 1, 0), child(n_nodes > 1 ? n_nodes - 1 : 0);
  std::vector<unsigned> word(n_nodes, 0);
  int nWords = 0;
  for (int i = 1; i < n_nodes; ++i) {
    if (parent[i] < 0 || parent[i] >= i) return orbfe_fail(ORBFE_ERR_INVALID, "node %d: parent %d does not precede it", i, parent[i]);
    cnt[parent[i] + 1]++;
    if (is_leaf[i]) word[i] = (unsigned)nWords++;
  }
  for (int i = 0; i < n_nodes; ++i) cnt[i + 1] += cnt[i];
  {
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int i = 1; i < n_nodes; ++i) child[fill[parent[i]]++] = i;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
  if (device < 0 || device >= ndev)
    return orbfe_fail(ORBFE_ERR_CUDA, "CUDA device %d not available (%d visible); this library has no CPU path", device, ndev);
  orbfe_vocabulary* v = new (std::nothrow) orbfe_vocabulary();
  if (!v) return orbfe_fail(ORBFE_ERR_NOMEM, "out of host memory");
  v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nNodes = n_nodes; v->nWords = nWords;
  std::vector<double> w0(1, 0.0);
  std::vector<uint8_t> d0(32, 0);
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&v->d_desc, (size_t)n_nodes * 32);
  if (e == cudaSuccess) e = cudaMalloc(&v->d_childStart, ((size_t)n_nodes + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&v->d_child, (child.size() + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&v->d_weight, (size_t)n_nodes * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&v->d_wordId, (size_t)n_nodes * sizeof(unsigned));
  // node 0 (root): no descriptor / weight in the file
  if (e == cudaSuccess) e = cudaMemcpy(v->d_desc, d0.data(), 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->d_weight, w0.data(), sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n_nodes > 1) e = cudaMemcpy(v->d_desc + 32, desc + 32, ((size_t)n_nodes - 1) * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && n_nodes > 1) e = cudaMemcpy(v->d_weight + 1, weight + 1, ((size_t)n_nodes - 1) * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->d_childStart, cnt.data(), ((size_t)n_nodes + 1) * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !child.empty()) e = cudaMemcpy(v->d_child, child.data(), child.size() * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(v->d_wordId, word.data(), (size_t)n_nodes * sizeof(unsigned), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    orbfe_vocabulary_destroy(v);
    return orbfe_fail(ORBFE_ERR_CUDA, "vocabulary upload failed: %s", cudaGetErrorString(e));
  }
  *out = v;
  return ORBFE_OK;
}

// TemplatedVocabulary::loadFromTextFile (TemplatedVocabulary.h:1335-1422): "k L scoring weighting" then one line per
// node "parent is_leaf d0 .. d31 weight"
int orbfe_vocabulary_load_text(const char* path, int device, orbfe_vocabulary** out) {
  if (!path || !out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *out = nullptr;
  std::ifstream f(path);
  if (!f.is_open()) return orbfe_fail(ORBFE_ERR_INVALID, "cannot open vocabulary file %s", path);
  std::string s;
  std::getline(f, s);
  std::stringstream ss(s);
  int k = -1, L = -1, n1 = -1, n2 = -1;
  ss >> k >> L >> n1 >> n2;
  if (ss.fail() || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3)
    return orbfe_fail(ORBFE_ERR_INVALID, "%s is not a vocabulary text file", path);
  std::vector<int32_t> parent(1, 0);
  std::vector<uint8_t> leaf(1, 0), desc(32, 0);
  std::vector<double> weight(1, 0.0);
  while (std::getline(f, s)) {
    if (s.find_first_not_of(" \t\r\n") == std::string::npos) continue;  // the reference would append a broken node for a blank tail line
    std::stringstream sn(s);
    int pid = 0, isLeaf = 0;
    sn >> pid >> isLeaf;
    uint8_t d[32];
    for (int i = 0; i < 32; ++i) {  // FORB::fromString (FORB.cpp:120-135)
      int b = 0;
      sn >> b;
      d[i] = (uint8_t)b;
    }
    double w = 0;
    sn >> w;
    if (sn.fail()) return orbfe_fail(ORBFE_ERR_INVALID, "%s: malformed node line %zu", path, parent.size());
    parent.push_back(pid); leaf.push_back(isLeaf > 0); weight.push_back(w);
    desc.insert(desc.end(), d, d + 32);
  }
  return orbfe_vocabulary_create(device, k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), out);
}

int orbfe_vocabulary_info(const orbfe_vocabulary* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words) {
  if (!v) return orbfe_fail(ORBFE_ERR_INVALID, "null vocabulary");
  if (k) *k = v->k;
  if (L) *L = v->L;
  if (scoring) *scoring = v->scoring;
  if (weighting) *weighting = v->weighting;
  if (n_nodes) *n_nodes = v->nNodes;
  if (n_words) *n_words = v->nWords;
  return ORBFE_OK;
}

int orbfe_bow_transform(orbfe_vocabulary* v, int n, const uint8_t* desc, int levelsup, uint32_t* word_id, uint32_t* node_id,
                        uint32_t* bow_words, double* bow_values, int* n_bow, uint32_t* fv_nodes, int32_t* fv_start,
                        uint32_t* fv_idx, int* n_fv) {
  if (!v || n < 0 || (n && !desc) || !n_bow || !n_fv) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n && (!bow_words || !bow_values || !fv_nodes || !fv_start || !fv_idx)) return orbfe_fail(ORBFE_ERR_INVALID, "null output array");
  *n_bow = 0; *n_fv = 0;
  if (fv_start) fv_start[0] = 0;
  if (n == 0 || v->nWords == 0 || v->nNodes < 2) return ORBFE_OK;  // empty() (:1132)
  CUDA_TRY(cudaSetDevice(v->device));
  int rc;
  if ((rc = ensure_scratch(v, n))) return rc;
  int n2 = 1024;
  while (n2 < n) n2 <<= 1;
  cudaStream_t st = v->stream;
  // output block for n features: counts | bowValue (f64) | word | node | bowWord | fvNode | fvIdx | fvStart (n + 1)
  const size_t N = ((size_t)n + 3) & ~(size_t)3;
  size_t off = 16;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  const size_t oVal = take(N * 8), oWord = take(N * 4), oNode = take(N * 4), oBw = take(N * 4), oFn = take(N * 4), oFi = take(N * 4),
               oFs = take((N + 4) * 4);
  v->d_counts = reinterpret_cast<int*>(v->d_out);
  v->d_bowValue = reinterpret_cast<double*>(v->d_out + oVal);
  v->d_word = reinterpret_cast<unsigned*>(v->d_out + oWord); v->d_node = reinterpret_cast<unsigned*>(v->d_out + oNode);
  v->d_bowWord = reinterpret_cast<unsigned*>(v->d_out + oBw); v->d_fvNode = reinterpret_cast<unsigned*>(v->d_out + oFn);
  v->d_fvIdx = reinterpret_cast<unsigned*>(v->d_out + oFi); v->d_fvStart = reinterpret_cast<int*>(v->d_out + oFs);
  memcpy(v->h_feat, desc, (size_t)n * 32);
  CUDA_TRY(cudaMemcpyAsync(v->d_feat, v->h_feat, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  BOW_LAUNCH(v, k_bow_descend, dim3((n + ORBFE_BOW_THREADS / 32 - 1) / (ORBFE_BOW_THREADS / 32)), dim3(ORBFE_BOW_THREADS), v->tree(),
             v->d_feat, n, levelsup, v->d_word, v->d_node, v->d_fw, v->d_wkey, v->d_nkey);
  // scoring -> normalisation (ScoringObject.h:73-90): L2_NORM -> L2, DOT_PRODUCT -> none, everything else L1
  const int norm = v->scoring == 1 ? 2 : (v->scoring == 5 ? 0 : 1);
  BowOut O;
  O.bowWord = v->d_bowWord; O.bowValue = v->d_bowValue; O.bowStart = v->d_bowStart; O.fvNode = v->d_fvNode;
  O.fvStart = v->d_fvStart; O.fvIdx = v->d_fvIdx; O.counts = v->d_counts;
  BOW_LAUNCH(v, k_bow_assemble, dim3(1), dim3(1024), v->d_wkey, v->d_nkey, n, n2, v->d_fw, v->weighting, norm, O);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(v->h_out, v->d_out, off, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const int* hc = reinterpret_cast<const int*>(v->h_out);
  const int nb = hc[0], nf = hc[1], kept = hc[2];
  if (nb < 0 || nb > n || nf < 0 || nf > n || kept < 0 || kept > n) return orbfe_fail(ORBFE_ERR_CUDA, "vocabulary transform returned inconsistent counts");
  if (word_id) memcpy(word_id, v->h_out + oWord, (size_t)n * sizeof(unsigned));
  if (node_id) memcpy(node_id, v->h_out + oNode, (size_t)n * sizeof(unsigned));
  if (nb) {
    memcpy(bow_words, v->h_out + oBw, (size_t)nb * sizeof(unsigned));
    memcpy(bow_values, v->h_out + oVal, (size_t)nb * sizeof(double));
  }
  if (nf) {
    memcpy(fv_nodes, v->h_out + oFn, (size_t)nf * sizeof(unsigned));
    memcpy(fv_idx, v->h_out + oFi, (size_t)kept * sizeof(unsigned));
  }
  memcpy(fv_start, v->h_out + oFs, ((size_t)nf + 1) * sizeof(int));
  *n_bow = nb; *n_fv = nf;
  return ORBFE_OK;
}

}  // extern "C"
