// k_bow.cuh -- N3: Frame::ComputeBoW (frame.cpp:258-263) = DBoW2 TemplatedVocabulary::transform
// (third_party/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1250) with FORB::distance (FORB.cpp:81-101).
//
//   k_bow_descend   one warp per descriptor: walks the vocabulary tree from the root; at every node the lanes take
//                   the children (k <= 20 in ORBvoc, any k here), 8 x __popc on the 32-byte rows, shuffle arg-min on
//                   (distance, child position) == the reference's strict '<' scan in child order (:1224-1236).
//                   Emits the word id, the node id `levelsup` levels above the leaves, and two 64-bit sort keys.
//   k_bow_assemble  one CTA: the two std::map containers the reference fills feature by feature.
//                   BowVector (BowVector.cpp:34-47): bitonic sort of (word, feature) keys, one entry per distinct word,
//                   value = the word weight added once per occurrence IN A LOOP (the reference's `+=` sequence, so the
//                   double rounding is the same), then BowVector::normalize (:63-87) with its ascending-word serial sum.
//                   FeatureVector (FeatureVector.cpp:32-47): sort of (node, feature) keys -> node ids ascending, feature
//                   indices ascending inside a node (push_back order).
#pragma once
#include "orbfe_common.cuh"

struct VocabTree {
  const uint8_t* desc;     // nNodes x 32 (node 0 = root, unused)
  const int* childStart;   // nNodes + 1: CSR over `child`
  const int* child;        // children of a node in ascending node id (= push_back order, TemplatedVocabulary.h:1388)
  const double* weight;    // Node::weight
  const unsigned* wordId;  // Node::word_id (0 for nodes that are not flagged as leaves)
  int nNodes, L;
};

#define ORBFE_BOW_THREADS 128
#define ORBFE_BOW_NOKEY 0xffffffffffffffffull

__global__ void __launch_bounds__(ORBFE_BOW_THREADS)
k_bow_descend(const VocabTree V, const uint8_t* __restrict__ desc, const int n, const int levelsup,
              unsigned* __restrict__ wordOut, unsigned* __restrict__ nodeOut, double* __restrict__ weightOut,
              unsigned long long* __restrict__ wkey, unsigned long long* __restrict__ nkey) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (ORBFE_BOW_THREADS / 32) + (threadIdx.x >> 5);
  if (i >= n) return;
  const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)i * 32));
  const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)i * 32) + 1);
  const int nidLevel = V.L - levelsup;  // :1209
  unsigned nid = 0;                     // root when nid_level <= 0 (:1210); also the value kept when a leaf is reached
                                        // above nid_level, where the reference leaves *nid unwritten (indeterminate)
  int finalId = 0, level = 0;
  int cs = V.childStart[0], ce = V.childStart[1];
  do {
    ++level;
    unsigned best = 0xffffffffu;  // distance << 16 | child position
    for (int c = cs + lane; c < ce; c += 32) {
      const int id = V.child[c];
      const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(V.desc + (size_t)id * 32));
      const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(V.desc + (size_t)id * 32) + 1);
      const unsigned d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
      best = min(best, (d << 16) | (unsigned)(c - cs));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    finalId = V.child[cs + (int)(best & 0xffffu)];
    if (level == nidLevel) nid = (unsigned)finalId;  // :1240-1241
    cs = V.childStart[finalId];
    ce = V.childStart[finalId + 1];
  } while (ce > cs);  // !isLeaf() == has children (:1243)
  if (lane == 0) {
    const unsigned word = V.wordId[finalId];
    const double w = V.weight[finalId];
    wordOut[i] = word;
    nodeOut[i] = nid;
    weightOut[i] = w;
    const bool keep = w > 0;  // not stopped (:1155)
    wkey[i] = keep ? ((unsigned long long)word << 32) | (unsigned)i : ORBFE_BOW_NOKEY;
    nkey[i] = keep ? ((unsigned long long)nid << 32) | (unsigned)i : ORBFE_BOW_NOKEY;
  }
}

// ascending bitonic sort of n2 (power of two) keys anywhere in memory, whole CTA
__device__ __forceinline__ void orbfe_block_sort_asc(unsigned long long* s, const int n2) {
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          const bool asc = (i & k) == 0;
          if (asc ? (a > b) : (a < b)) { s[i] = b; s[ixj] = a; }
        }
      }
    }
  __syncthreads();
}

struct BowOut {
  unsigned* bowWord;    // distinct words, ascending
  double* bowValue;
  int* bowStart;        // scratch: first sorted position of each distinct word (+ end)
  unsigned* fvNode;     // distinct nodes, ascending
  int* fvStart;         // CSR into fvIdx (+ end)
  unsigned* fvIdx;      // feature indices
  int* counts;          // [0] distinct words, [1] distinct nodes, [2] kept features
};

// weighting: 0 TF_IDF, 1 TF, 2 IDF, 3 BINARY (BowVector.h:36-42); norm: 0 none, 1 L1, 2 L2 (ScoringObject.h:73-90)
__global__ void __launch_bounds__(1024)
k_bow_assemble(unsigned long long* __restrict__ wkey, unsigned long long* __restrict__ nkey, const int n, const int n2,
               const double* __restrict__ featWeight, const int weighting, const int norm, const BowOut O) {
  __shared__ int s_scan[33];
  __shared__ double s_norm;
  const int tid = threadIdx.x, T = blockDim.x;
  for (int i = n + tid; i < n2; i += T) { wkey[i] = ORBFE_BOW_NOKEY; nkey[i] = ORBFE_BOW_NOKEY; }
  orbfe_block_sort_asc(wkey, n2);
  orbfe_block_sort_asc(nkey, n2);
  // ---- distinct heads: each thread owns a contiguous chunk of the sorted arrays
  const int per = (n2 + T - 1) / T;
  const int p0 = min(tid * per, n2), p1 = min(p0 + per, n2);
  for (int pass = 0; pass < 2; ++pass) {
    const unsigned long long* key = pass == 0 ? wkey : nkey;
    int heads = 0, kept = 0;
    for (int p = p0; p < p1; ++p) {
      const unsigned long long k = key[p];
      if (k == ORBFE_BOW_NOKEY) break;
      ++kept;
      if (p == 0 || (unsigned)(key[p - 1] >> 32) != (unsigned)(k >> 32)) ++heads;
    }
    int total;
    int u = orbfe_block_exscan(heads, s_scan, &total);
    int totalKept;
    orbfe_block_exscan(kept, s_scan, &totalKept);
    for (int p = p0; p < p1; ++p) {
      const unsigned long long k = key[p];
      if (k == ORBFE_BOW_NOKEY) break;
      if (p == 0 || (unsigned)(key[p - 1] >> 32) != (unsigned)(k >> 32)) {
        if (pass == 0) { O.bowWord[u] = (unsigned)(k >> 32); O.bowStart[u] = p; }
        else { O.fvNode[u] = (unsigned)(k >> 32); O.fvStart[u] = p; }
        ++u;
      }
      if (pass == 1) O.fvIdx[p] = (unsigned)k;
    }
    if (tid == 0) {
      O.counts[pass] = total;
      if (pass == 0) { O.bowStart[total] = totalKept; O.counts[2] = totalKept; }
      else O.fvStart[total] = totalKept;
    }
    __syncthreads();
  }
  // ---- BowVector values (TemplatedVocabulary.h:1141-1187)
  const int nU = O.counts[0];
  const bool accumulate = weighting == 0 || weighting == 1;  // addWeight vs addIfNotExist
  for (int u = tid; u < nU; u += T) {
    const int b = O.bowStart[u], e = O.bowStart[u + 1];
    const double w = featWeight[(unsigned)wkey[b]];  // the weight of this word (same for every feature of the run)
    double v = w;
    if (accumulate)
      for (int k = b + 1; k < e; ++k) v = __dadd_rn(v, w);  // vit->second += v, once per occurrence (BowVector.cpp:40)
    if (accumulate && norm == 0) v = __ddiv_rn(v, (double)nU);  // :1162-1168 (only when the scoring does not normalise)
    O.bowValue[u] = v;
  }
  __syncthreads();
  if (norm != 0) {  // BowVector::normalize (BowVector.cpp:63-87): serial sum in ascending word order
    if (tid == 0) {
      double acc = 0.0;
      if (norm == 1)
        for (int u = 0; u < nU; ++u) acc = __dadd_rn(acc, fabs(O.bowValue[u]));
      else {
        for (int u = 0; u < nU; ++u) acc = __dadd_rn(acc, __dmul_rn(O.bowValue[u], O.bowValue[u]));
        acc = __dsqrt_rn(acc);
      }
      s_norm = acc;
    }
    __syncthreads();
    const double nrm = s_norm;
    if (nrm > 0.0)
      for (int u = tid; u < nU; u += T) O.bowValue[u] = __ddiv_rn(O.bowValue[u], nrm);
  }
}
