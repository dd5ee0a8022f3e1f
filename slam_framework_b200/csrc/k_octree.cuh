// k_octree.cuh -- E4: quad-tree ("OctTree") keypoint distribution as a per-level GPU reduction
// that reproduces ORBextractor::DistributeOctTree + ExtractorNode::DivideNode
// (orb_extractor.cpp:422-704) exactly, including the node-splitting ORDER.
//
// One CTA per (level, image slot).  The reference's std::list algorithm is restated
// level-synchronously:
//   * a candidate only needs to know its owning node (knode[]); node membership is refined with
//     per-child atomic counters, no key permutation is needed because the final pick inside a
//     node is "max response, first in original order on ties" (:682-701) = max over
//     (response, -index);
//   * the std::list order is a function of creation order: every pass push_front()s the new
//     children, so the final list is "nodes by creation sequence, descending" with the initial
//     nodes (push_back) given sequence nIni-1-i.  Each node carries its creation sequence;
//     the finals are sorted by it once at the end (bitonic sort in shared memory);
//   * phase-1 passes (:535-606) split every expandable node, walking the list front to back
//     = previous generation in REVERSE creation order (initial nodes: forward order);
//   * careful rounds (:617-678) sort the previous generation by (size, creation sequence)
//     ascending (the reference's pointer tie-break is defined as creation sequence),
//     process from the back and stop as soon as the list holds >= N nodes; because each split
//     adds (#non-empty children - 1) >= 0 nodes, the stop index is a prefix scan.
// Expandable nodes live for exactly one generation, so two ping-pong generation buffers of
// capacity 4*max(N,nIni) and a finals list of capacity max(N+3,4*nIni) bound the memory.
#pragma once
#include "orbfe_common.cuh"

#ifndef ORBFE_OCT_THREADS
#define ORBFE_OCT_THREADS 512  // A/B on B200, 64 pairs (threads / min CTAs per SM / smem budget): 512/4/50 KB 0.119 ms, 256/8/25 KB 0.129, 256/6/34 KB 0.142, 1024/2/100 KB 0.184; a single pair: 512 threads 55.1 us, 1024 threads 57.6 us
#endif
#define ORBFE_OCT_NEW 0x40000000

struct OctNode {
  short x0, x1, y0, y1;
  int cnt;
  int seq;
};

struct OctScratch {      // per-slot strided device arrays (bases for slot 0)
  unsigned* cand;        // candStride per slot
  int* knode;            // candStride per slot
  int* cellStart;        // totalCells per slot
  OctNode* nodes;        // 2 * nodeStride per slot (two generations, level blocks inside)
  int* childCnt;         // nodeStride per slot (4 counters per node of the current generation)
  int* childSlot;        // nodeStride per slot
  unsigned long long* best;  // nodeStride*5/4+16 per slot
  int* finSeq;           // totalOut per slot
  int* finKey;           // totalOut per slot
};

#ifndef ORBFE_OCT_MINB
#define ORBFE_OCT_MINB 4  // <= 32 registers: 4 CTAs (64 warps) per SM for this latency-bound kernel (A/B measured)
#endif
__global__ void __launch_bounds__(ORBFE_OCT_THREADS, ORBFE_OCT_MINB)
k_octree(const __grid_constant__ Geom g, const int* __restrict__ cellCnt, const unsigned* __restrict__ cellList,
         const OctScratch sc, unsigned* __restrict__ lvlKp, int* __restrict__ lvlCnt, int* __restrict__ errFlag,
         const int stageCap) {
  ORBFE_DYN_SMEM(smem);
  unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(smem);  // g.sortCap entries
  __shared__ int s_scan[33];
  __shared__ int s_i[16];
  const int T = ORBFE_OCT_THREADS;
  const int tid = threadIdx.x;
  // grid = (slots, levels): blocks are dispatched x-first, so the long level-0 trees of every image start first and the
  // short top-level ones fill the tail of the launch
  const int level = blockIdx.y, slot = blockIdx.x;
  const bool latencyForm = gridDim.x <= 8;  // a few frames in flight: barrier-bound, see orbfe_block_sort_desc
  const LevelGeom& L = g.lv[level];
  const int N = L.N;

  const int* cc = cellCnt + (size_t)slot * g.totalCells + L.cellBase;
  const unsigned* cl = cellList + (size_t)slot * g.cellListStride + L.cellListOff;
  unsigned* cand = sc.cand + (size_t)slot * g.candStride + L.candOff;
  int* knode = sc.knode + (size_t)slot * g.candStride + L.candOff;
  // the key arrays are walked three times per generation with dependent loads: when the level's candidates fit
  // (stageCap entries, the common case) they live in shared memory instead of L2
  unsigned* const candG = cand;  // still written: orbfe_debug_candidates (per-stage parity tap) reads it
  int* cellStart = sc.cellStart + (size_t)slot * g.totalCells + L.cellBase;
  OctNode* genA = sc.nodes + (size_t)slot * 2 * g.nodeStride + 2 * (size_t)L.nodeOff;
  OctNode* genB = genA + L.nodeCap;
  // per-child key counters and child slots live in shared memory: thousands of keys hit a handful of
  // counters in the first passes, which serialises global (L2) atomics for tens of microseconds
  int* childCnt = reinterpret_cast<int*>(s_sort + g.sortCap);
  int* childSlot = childCnt + g.maxNodeCap;
  unsigned long long* best = sc.best + (size_t)slot * (g.nodeStride * 5 / 4 + 16 * ORBFE_MAX_LEVELS) + (size_t)L.nodeOff * 5 / 4 + 16 * level;
  int* finSeq = sc.finSeq + (size_t)slot * g.totalOut + L.outOff;
  int* finKey = sc.finKey + (size_t)slot * g.totalOut + L.outOff;
  unsigned* out = lvlKp + (size_t)slot * g.totalOut + L.outOff;
  int* outCnt = lvlCnt + (size_t)slot * g.nlevels + level;

  // ---- 1. gather the per-cell lists into the reference's vToDistributeKeys order --------
  const int nCells = L.nCols * L.nRows;
  int n = 0;
  {
    const int per = (nCells + T - 1) / T;
    const int c0 = min(tid * per, nCells), c1 = min(c0 + per, nCells);
    int sum = 0;
    for (int c = c0; c < c1; ++c) sum += cc[c];
    int total;
    int run = orbfe_block_exscan(sum, s_scan, &total);
    for (int c = c0; c < c1; ++c) { cellStart[c] = run; run += cc[c]; }
    n = total;
    if (n <= stageCap) {
      cand = reinterpret_cast<unsigned*>(childSlot + g.maxNodeCap);
      knode = reinterpret_cast<int*>(cand + stageCap);
    }
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5;
    for (int c = wid; c < nCells; c += T / 32) {
      const int cn = cc[c], st = cellStart[c];
      for (int k = lane; k < cn; k += 32) {
        const unsigned v = cl[(size_t)c * L.cellCap + k];
        cand[st + k] = v;
        if (cand != candG) candG[st + k] = v;
      }
    }
    __syncthreads();
  }
  if (n == 0) {
    if (tid == 0) *outCnt = 0;
    return;
  }

  // ---- 2. initial nodes (:484-531) ----------------------------------------------------------
  const int nIni = L.nIni;
  const float hX = L.hX;
  for (int r = tid; r < 4 * nIni; r += T) childCnt[r] = 0;
  __syncthreads();
  for (int k = tid; k < n; k += T) {
    int r = (int)__fdiv_rn((float)ORBFE_PX(cand[k]), hX);
    if (r >= nIni) r = nIni - 1;
    knode[k] = r;
    atomicAdd(&childCnt[r], 1);
  }
  __syncthreads();
  // thread 0 lays out the roots: expandable roots are stored so that REVERSE storage order is
  // forward root order; single-key roots become finals with seq nIni-1-r.
  if (tid == 0) {
    int m = 0, nf = 0, alive = 0;
    for (int r = 0; r < nIni; ++r) m += childCnt[r] > 1;
    int j = m;
    for (int r = 0; r < nIni; ++r) {
      const int c = childCnt[r];
      if (c == 0) { childSlot[r] = -1; continue; }
      ++alive;
      if (c == 1) { finSeq[nf] = nIni - 1 - r; childSlot[r] = -(nf + 2); ++nf; continue; }
      --j;
      OctNode nd;
      nd.x0 = (short)(int)__fmul_rn(hX, (float)r);
      nd.x1 = (short)(int)__fmul_rn(hX, (float)(r + 1));
      nd.y0 = 0; nd.y1 = (short)L.boxH;
      nd.cnt = c; nd.seq = nIni - 1 - r;
      genA[j] = nd;
      childSlot[r] = j;
    }
    s_i[0] = m;       // size of the current generation
    s_i[1] = nf;      // finals so far
    s_i[2] = alive;   // lNodes.size()
    s_i[3] = nIni;    // next creation sequence
    s_i[4] = 0;       // mode: 0 = phase-1 passes, 1 = careful rounds
    s_i[5] = 0;       // finished
  }
  __syncthreads();
  for (int k = tid; k < n; k += T) {
    const int s = childSlot[knode[k]];
    if (s >= 0) knode[k] = s << 2;
    else { finKey[-(s + 2)] = k; knode[k] = -1; }
  }
  __syncthreads();

  OctNode* cur = genA;
  OctNode* nxt = genB;
  // ---- 3. passes ------------------------------------------------------------------------------
  for (int iter = 0; iter < 64; ++iter) {
    const int m = s_i[0], nfin = s_i[1], prevSize = s_i[2], seq0 = s_i[3], mode = s_i[4];
    if (m == 0) break;  // every node holds a single key: list cannot change (:610)
    for (int t = tid; t < 4 * m; t += T) childCnt[t] = 0;
    __syncthreads();
    // B: count keys per child
    for (int k = tid; k < n; k += T) {
      const int v = knode[k];
      if (v < 0) continue;
      const int j = (v & ~ORBFE_OCT_NEW) >> 2;
      const OctNode nd = cur[j];
      const unsigned c = cand[k];
      const int midX = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1);  // ceil((UR.x-UL.x)/2.f), :424
      const int midY = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
      const int q = (ORBFE_PX(c) < midX ? 0 : 1) + (ORBFE_PY(c) < midY ? 0 : 2);
      knode[k] = (j << 2) | q;
      atomicAdd(&childCnt[4 * j + q], 1);
    }
    __syncthreads();
    // D: processing order -> s_sort[p] low 32 bits = node index
    int m2 = 1;
    if (mode == 0) {
      for (int p = tid; p < m; p += T) s_sort[p] = (unsigned long long)(m - 1 - p);
      __syncthreads();
    } else {
      while (m2 < m) m2 <<= 1;
      for (int p = tid; p < m2; p += T)
        s_sort[p] = p < m ? (((unsigned long long)(unsigned)cur[p].cnt << 32) | (unsigned)p) : 0ull;
      // (cnt, creation index) descending == back-to-front walk of the ascending std::sort (:625-627)
      orbfe_block_sort_desc(s_sort, m2, latencyForm);
    }
    // E: prefix scans in processing order (chunks of T)
    int baseNe = 0, baseNx = 0, baseDelta = 0, nSplit = 0;
    for (int p0 = 0; p0 < m; p0 += T) {
      const int p = p0 + tid;
      int ne = 0, nx = 0;
      int j = -1;
      if (p < m) {
        j = (int)(s_sort[p] & 0xffffffffu);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int c = childCnt[4 * j + q]; ne += c > 0; nx += c > 1; }
      }
      int totDelta, totNe, totNx;
      const int exDelta = orbfe_block_exscan(p < m ? ne - 1 : 0, s_scan, &totDelta);
      // a node is split iff the list was still below N before it (:672-673); phase-1 splits all
      const bool split = p < m && (mode == 0 || prevSize + baseDelta + exDelta < N);
      int exNe, exNx;
      if (L.nodeCap < 65536) {  // block-uniform: the two child counts share one scan while their sums fit 16 bits each
        int totP;               // (a generation creates at most nodeCap children)
        const int exP = orbfe_block_exscan(split ? (ne | (nx << 16)) : 0, s_scan, &totP);
        exNe = exP & 0xffff; exNx = exP >> 16;
        totNe = totP & 0xffff; totNx = totP >> 16;
      } else {
        exNe = orbfe_block_exscan(split ? ne : 0, s_scan, &totNe);
        exNx = orbfe_block_exscan(split ? nx : 0, s_scan, &totNx);
      }
      if (p < m) {
        if (split) {
          const OctNode nd = cur[j];
          const int midX = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1);
          const int midY = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
          int ie = 0, ix = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = childCnt[4 * j + q];
            if (c == 0) { childSlot[4 * j + q] = -1; continue; }
            const int seq = seq0 + baseNe + exNe + ie;
            if (c == 1) {
              const int f = nfin + (baseNe - baseNx) + (exNe - exNx) + (ie - ix);
              finSeq[f] = seq;
              childSlot[4 * j + q] = -(f + 2);
            } else {
              const int s = baseNx + exNx + ix;
              OctNode ch;
              ch.x0 = (q & 1) ? (short)midX : nd.x0;
              ch.x1 = (q & 1) ? nd.x1 : (short)midX;
              ch.y0 = (q & 2) ? (short)midY : nd.y0;
              ch.y1 = (q & 2) ? nd.y1 : (short)midY;
              ch.cnt = c; ch.seq = seq;
              nxt[s] = ch;
              childSlot[4 * j + q] = s;
              ++ix;
            }
            ++ie;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) childSlot[4 * j + q] = -3 - 0x10000000;  // marker: parent not split
        }
      }
      baseDelta += totDelta; baseNe += totNe; baseNx += totNx;
      nSplit += __syncthreads_count(split);
    }
    __syncthreads();
    // G: move keys to their child
    for (int k = tid; k < n; k += T) {
      const int v = knode[k];
      if (v < 0 || (v & ORBFE_OCT_NEW)) continue;
      const int s = childSlot[v];  // v = 4*j+q
      if (s >= 0) knode[k] = ORBFE_OCT_NEW | (s << 2);
      else if (s == -3 - 0x10000000) { /* parent survives un-split (only when finishing) */ }
      else { finKey[-(s + 2)] = k; knode[k] = -1; }
    }
    __syncthreads();
    // H: bookkeeping + termination (:608-614, :680)
    const int newSize = prevSize - nSplit + baseNe;
    const bool finish = newSize >= N || newSize == prevSize;
    if (tid == 0) {
      s_i[1] = nfin + (baseNe - baseNx);
      s_i[2] = newSize;
      s_i[3] = seq0 + baseNe;
      s_i[6] = m;        // previous generation size (for the un-split survivors)
      s_i[7] = baseNx;   // new generation size
      if (finish) s_i[5] = 1;
      else {
        if (mode == 0 && newSize + 3 * baseNx > N) s_i[4] = 1;
        s_i[0] = baseNx;
      }
    }
    __syncthreads();
    if (finish) break;
    // next generation becomes current.  The NEW flag of the moved keys needs no clearing pass: step B of the next generation
    // masks it when it reads a key and rewrites the key without it, and nothing reads knode[] in between (the barrier after
    // the counter reset below orders G's writes before B's reads).
    OctNode* t = cur; cur = nxt; nxt = t;
  }

  // ---- 4. finals: add the expandable nodes still alive, pick the best key of each (:682-701) --
  const int finished = s_i[5];
  int nfin = s_i[1];
  const int listSize = s_i[2];
  int mOld = 0, mNew = 0;
  if (finished) { mOld = s_i[6]; mNew = s_i[7]; } else { mOld = 0; mNew = 0; }
  // (not finished only when the loop ended with m == 0: all nodes are finals already)
  for (int t = tid; t < mOld + mNew; t += T) best[t] = 0ull;
  __syncthreads();
  for (int k = tid; k < n; k += T) {
    const int v = knode[k];
    if (v < 0) continue;
    const int idx = (v & ORBFE_OCT_NEW) ? (mOld + ((v & ~ORBFE_OCT_NEW) >> 2)) : (v >> 2);
    const unsigned long long key = ((unsigned long long)ORBFE_PS(cand[k]) << 32) | (unsigned)(0x7fffffff - k);
    atomicMax(&best[idx], key);
  }
  __syncthreads();
  {
    // alive expandable nodes = un-split old nodes (best != 0) and every new node
    int totA = 0;
    for (int t0 = 0; t0 < mOld + mNew; t0 += T) {
      const int t = t0 + tid;
      const bool alive = t < mOld + mNew && best[t] != 0ull;
      int tot;
      const int ex = orbfe_block_exscan(alive ? 1 : 0, s_scan, &tot);
      if (alive) {
        const int f = nfin + totA + ex;
        if (f < L.outCap) {
          finSeq[f] = t < mOld ? cur[t].seq : nxt[t - mOld].seq;
          finKey[f] = 0x7fffffff - (int)(best[t] & 0xffffffffu);
        }
      }
      totA += tot;
    }
    nfin += totA;
  }
  __syncthreads();
  if (nfin != listSize || nfin > L.outCap) {
    if (tid == 0) { atomicOr(errFlag, 1); *outCnt = 0; }
    return;
  }
  // ---- 5. list order = creation sequence descending ------------------------------------------
  int n2 = 1;
  while (n2 < nfin) n2 <<= 1;
  for (int t = tid; t < n2; t += T)
    s_sort[t] = t < nfin ? ((((unsigned long long)(unsigned)finSeq[t] + 1ull) << 32) | (unsigned)t) : 0ull;
  orbfe_block_sort_desc(s_sort, n2, latencyForm);
  for (int t = tid; t < nfin; t += T) out[t] = cand[finKey[(int)(s_sort[t] & 0xffffffffu)]];
  if (tid == 0) *outCnt = nfin;
}
