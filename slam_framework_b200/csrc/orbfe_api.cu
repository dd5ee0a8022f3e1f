// orbfe_api.cu -- host side of liborbfe.so: the C ABI of include/orbfe.h over the sm_100a kernels
// (k_pyramid / k_fast / k_octree / k_blur / k_describe / k_stereo).  No CPU fallback: every compute entry
// point needs a CUDA device and fails with ORBFE_ERR_CUDA otherwise.
//
// One handle = one device arena sized for `max_images` image slots + one private stream.  A batch
// of n images is processed by 8 pyramid launches (the level chain of orb_extractor.cpp:1051-1076 is
// sequential by definition) + 4 launches (FAST, quad-tree, blur, orientation+descriptor), each
// launch covering every (slot, level, tile) of the batch; stereo adds 2 launches per batch of pairs.
#include "../../include/orbfe.h"

#include "k_pyramid.cuh"
#include "k_fast.cuh"
#include "k_octree.cuh"
#include "k_blur.cuh"
#include "k_describe.cuh"
#include "k_stereo.cuh"
#include "k_gray.cuh"
#include "orbfe_host.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

static_assert(sizeof(orbfe_keypoint) == 28, "cv::KeyPoint layout");
static_assert(sizeof(orbfe_kp_dev) == 28, "cv::KeyPoint layout");

// ---- error plumbing ---------------------------------------------------------------------------
static thread_local std::string t_last_error;
int orbfe_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
static inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }
// cvRound: round-half-to-even in the default rounding mode
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }

struct orbfe_extractor {
  int device = 0;
  cudaStream_t stream = nullptr;
  orbfe_params p{};
  int S = 1;
  float scale[ORBFE_MAX_LEVELS], invScale[ORBFE_MAX_LEVELS], sigma2[ORBFE_MAX_LEVELS], invSigma2[ORBFE_MAX_LEVELS];
  int fpl[ORBFE_MAX_LEVELS];
  Geom g{};
  bool configured = false;
  int fastRows = 0, fastQ1Cap = 0, fastQ2Cap = 0, fastListCap = 0;  // FAST tile rows (TMA box height), queue / list capacities
  size_t fastSmem = 0, octSmem = 0;
  CUtensorMap* d_tmaps = nullptr;   // per level: the padded pyramid planes of every slot (box = 256 B x fastRows)
  unsigned* d_fastTasks = nullptr;  // per FAST CTA: level | cell row | first cell
  CUtensorMap* d_tmapsBlur = nullptr;  // the same planes with the blur's box (ORBFE_BLUR_BOXW x ORBFE_BLUR_RB)
  unsigned* d_blurTasks = nullptr;  // per blur CTA: level | column strip
  CUtensorMap* d_tmapsPyr = nullptr;   // per level l >= 1: the planes of level l-1 with the streaming resize's box
  CUtensorMap* d_tmapsDescPyr = nullptr;   // k_orient_describe: padded planes, box 48 B x 31 rows (the IC_Angle disc)
  CUtensorMap* d_tmapsDescBlur = nullptr;  // k_orient_describe: blurred planes, box 64 B x 37 rows (the rBRIEF window)
  int* d_pyrBoxX = nullptr;         // per (level, strip): first source byte of the strip's boxes
  int octStageCap = 0;
  // device arena
  uint8_t* d_img = nullptr;
  uint8_t* d_color = nullptr;       // staging of colour frames (orbfe_upload_color), allocated on first use
  size_t colorStride = 0;
  uint8_t* d_pyr = nullptr;
  uint8_t* d_blur = nullptr;
  int* d_cellCnt = nullptr;
  unsigned* d_cellList = nullptr;
  OctScratch oct{};
  size_t bestStride = 0;
  unsigned* d_lvlKp = nullptr;
  int* d_lvlCnt = nullptr;
  orbfe_kp_dev* d_kps = nullptr;
  uint8_t* d_desc = nullptr;
  int* d_nKp = nullptr;
  ResizeLut* d_lut = nullptr;
  PyrWordLut* d_wlut = nullptr;
  PyrRowLut* d_rlut = nullptr;
  int* d_err = nullptr;
  uint2* d_icw = nullptr;           // IC_Angle per-byte weight table (k_orient_describe)
  // stereo
  StereoPair* d_pairs = nullptr;
  float* d_uR = nullptr;
  float* d_depth = nullptr;
  int* d_sad = nullptr;
  int* d_nMatched = nullptr;
  int* d_rowStart = nullptr;        // per pair: h0 + 1 offsets
  uint2* d_rowItems = nullptr;      // per pair: rowCap (index | octave, x) entries
  int rowCap = 0;
  size_t rowSmem = 0;
  // pinned staging
  int* h_n = nullptr;          // S counts + S matched + 1 err + 2 scratch (orbfe_stereo_match)
  orbfe_kp_dev* h_kps = nullptr;
  uint8_t* h_desc = nullptr;
  float* h_uR = nullptr;
  float* h_depth = nullptr;
  StereoPair* h_pairs = nullptr;
  cudaEvent_t ev[64] = {};
  cudaEvent_t stageEv[64][8] = {};  // ring of per-run stage brackets
  unsigned char stageHas[64] = {};  // 1 = extract recorded, 2 = stereo recorded too
  int stageRuns = 0;                // runs recorded since the last summary
  int pairsCached = 0;              // d_pairs holds the same-handle table for this many pairs
  // single-frame / single-pair calls (orbfe_extract, orbfe_extract_batch with <= 2 images): the whole sequence
  // H2D -> 15 kernels -> D2H is captured ONCE per geometry into a CUDA graph and replayed per frame (one launch call)
  uint8_t* h_img = nullptr;         // pinned staging of <= 2 input frames (graph memcpy nodes need a fixed source)
#ifndef ORBFE_EMU
  cudaGraphExec_t xGraph[2][4] = {};  // [n_imgs - 1][kps | desc << 1]
  int xGraphLaunches[2][4] = {};
#endif
  bool stageTiming = false;
  bool slotDirty = true;            // the device keypoint rows are NOT what h_n / h_kps / h_desc hold (no download yet, or overwritten)
  long long launches = 0;
};

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return orbfe_fail(ORBFE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#ifdef ORBFE_EMU
#define ORBFE_LAUNCH(ex, kernel, grid, block, smem, ...)                                           \
  do { if (getenv("ORBFE_EMU_TRACE")) fprintf(stderr, "launch %s\n", #kernel);                      \
       emu::launch(grid, block, smem, [&]() { kernel(__VA_ARGS__); }); (ex)->launches++; } while (0)
#else
#define ORBFE_LAUNCH(ex, kernel, grid, block, smem, ...)                                           \
  do { kernel<<<grid, block, smem, (ex)->stream>>>(__VA_ARGS__); (ex)->launches++; } while (0)
#endif

// ---- geometry ---------------------------------------------------------------------------------
// cv::resize(INTER_LINEAR) coefficient tables for one axis
static void build_resize_lut(int dn, int sn, ResizeLut* out) {
  const double inv_scale = (double)dn / sn;
  const double scale = 1.0 / inv_scale;
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = cv_floor_f(f);
    f -= (float)s;
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    out[d].ofs = s;
    out[d].c0 = (short)cv_round_f((1.f - f) * 2048.f);
    out[d].c1 = (short)cv_round_f(f * 2048.f);
  }
}

static void free_arena(orbfe_extractor* ex) {
  cudaFree(ex->d_color); ex->d_color = nullptr; ex->colorStride = 0;
  cudaFree(ex->d_tmaps); ex->d_tmaps = nullptr; cudaFree(ex->d_fastTasks); ex->d_fastTasks = nullptr;
  cudaFree(ex->d_tmapsBlur); ex->d_tmapsBlur = nullptr; cudaFree(ex->d_blurTasks); ex->d_blurTasks = nullptr;
  cudaFree(ex->d_tmapsPyr); ex->d_tmapsPyr = nullptr; cudaFree(ex->d_pyrBoxX); ex->d_pyrBoxX = nullptr;
  cudaFree(ex->d_tmapsDescPyr); ex->d_tmapsDescPyr = nullptr; cudaFree(ex->d_tmapsDescBlur); ex->d_tmapsDescBlur = nullptr;
  cudaFree(ex->d_img); cudaFree(ex->d_pyr); cudaFree(ex->d_blur); cudaFree(ex->d_cellCnt); cudaFree(ex->d_cellList);
  cudaFree(ex->oct.cand); cudaFree(ex->oct.knode); cudaFree(ex->oct.cellStart); cudaFree(ex->oct.nodes);
  cudaFree(ex->oct.childCnt); cudaFree(ex->oct.childSlot); cudaFree(ex->oct.best); cudaFree(ex->oct.finSeq);
  cudaFree(ex->oct.finKey); cudaFree(ex->d_lvlKp); cudaFree(ex->d_lvlCnt); cudaFree(ex->d_kps); cudaFree(ex->d_desc);
  cudaFree(ex->d_icw); ex->d_icw = nullptr; cudaFree(ex->d_nKp); cudaFree(ex->d_lut); cudaFree(ex->d_wlut); ex->d_wlut = nullptr; cudaFree(ex->d_rlut); ex->d_rlut = nullptr; cudaFree(ex->d_err); cudaFree(ex->d_pairs); cudaFree(ex->d_uR);
  cudaFree(ex->d_depth); cudaFree(ex->d_sad); cudaFree(ex->d_nMatched); cudaFree(ex->d_rowStart); cudaFree(ex->d_rowItems);
  ex->d_rowStart = nullptr; ex->d_rowItems = nullptr;
  cudaFreeHost(ex->h_n); cudaFreeHost(ex->h_kps); cudaFreeHost(ex->h_desc); cudaFreeHost(ex->h_uR);
  cudaFreeHost(ex->h_depth); cudaFreeHost(ex->h_pairs);
#ifndef ORBFE_EMU
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 4; ++b)
      if (ex->xGraph[a][b]) { cudaGraphExecDestroy(ex->xGraph[a][b]); ex->xGraph[a][b] = nullptr; }
#endif
  cudaFreeHost(ex->h_img); ex->h_img = nullptr;
  ex->d_img = ex->d_pyr = ex->d_blur = nullptr; ex->d_cellCnt = nullptr; ex->d_cellList = nullptr;
  ex->oct = OctScratch{}; ex->d_lvlKp = nullptr; ex->d_lvlCnt = nullptr; ex->d_kps = nullptr; ex->d_desc = nullptr;
  ex->d_nKp = nullptr; ex->d_lut = nullptr; ex->d_err = nullptr; ex->d_pairs = nullptr; ex->d_uR = nullptr;
  ex->d_depth = nullptr; ex->d_sad = nullptr; ex->d_nMatched = nullptr;
  ex->h_n = nullptr; ex->h_kps = nullptr; ex->h_desc = nullptr; ex->h_uR = nullptr; ex->h_depth = nullptr;
  ex->h_pairs = nullptr;
  ex->pairsCached = 0;
  ex->configured = false;
}

// Computes the level geometry for a w0 x h0 input and (re)allocates the device arena.
static int configure(orbfe_extractor* ex, int w0, int h0) {
  if (ex->configured && ex->g.w0 == w0 && ex->g.h0 == h0) return ORBFE_OK;
  if (w0 > 4095 + 32 || h0 > 4095 + 32)
    return orbfe_fail(ORBFE_ERR_INVALID, "image %dx%d exceeds the 12-bit packed coordinate range", w0, h0);
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  free_arena(ex);
  Geom& g = ex->g;
  memset(&g, 0, sizeof(g));
  const int nl = ex->p.nlevels;
  g.nlevels = nl; g.w0 = w0; g.h0 = h0; g.iniTh = ex->p.ini_th_fast; g.minTh = ex->p.min_th_fast;
  // input frames are kept tightly packed (pitch == width) so that one contiguous H2D copy per image
  // (or per batch, when the host frames are contiguous) feeds them; k_pyramid_level re-aligns on the fly
  g.imgPitch = w0;
  g.imgStride = (unsigned)align_up_sz((size_t)w0 * h0, 4);
  size_t pyrOff = 0, blurOff = 0, cellListOff = 0, candOff = 0, nodeOff = 0;
  int cellBase = 0, outOff = 0, tileBase = 0, lutOff = 0, maxSort = 1;
  int maxInnerH = 1, fastBase = 0, maxQueue = 1, maxCellCap = 1;
  std::vector<unsigned> fastTasks, blurTasks;
  std::vector<int> pyrBoxX;
  std::vector<ResizeLut> lut;
  std::vector<PyrWordLut> wlut;
  std::vector<PyrRowLut> rlut;
  for (int l = 0; l < nl; ++l) {
    LevelGeom& L = g.lv[l];
    // level size (orb_extractor.cpp:1055-1056)
    L.w = cv_round_f((float)w0 * ex->invScale[l]);
    L.h = cv_round_f((float)h0 * ex->invScale[l]);
    if (L.w < 1 || L.h < 1) return orbfe_fail(ORBFE_ERR_INVALID, "pyramid level %d of a %dx%d image is empty", l, w0, h0);
    L.scale = ex->scale[l];
    L.kpSize = (float)(int)(ORBFE_PATCH * ex->scale[l]);  // :776 scaledPatchSize
    L.pitch = align_up(L.w + 2 * ORBFE_EDGE, 16);
    L.bpitch = align_up(L.w, 16);
    L.planeOff = (unsigned)pyrOff;
    pyrOff += align_up_sz((size_t)L.pitch * (L.h + 2 * ORBFE_EDGE), 256);
    L.blurOff = (unsigned)blurOff;
    blurOff += align_up_sz((size_t)L.bpitch * L.h, 256);
    // resize LUT (level l from level l-1)
    L.area2 = 0;
    if (l > 0) {
      const LevelGeom& P = g.lv[l - 1];
      L.area2 = (P.w == 2 * L.w && P.h == 2 * L.h) ? 1 : 0;  // exact 2x: OpenCV executes INTER_AREA
      L.lutXOff = lutOff; lutOff += L.w;
      L.lutYOff = lutOff; lutOff += L.h;
      lut.resize(lutOff);
      build_resize_lut(L.w, P.w, lut.data() + L.lutXOff);
      build_resize_lut(L.h, P.h, lut.data() + L.lutYOff);
    }
    L.pyrWords = (L.w + 2 * ORBFE_EDGE + 3) / 4;
    // per-word LUT of the fast resize kernel (k_pyramid_resize)
    L.wlutOff = (int)wlut.size();
    L.fastResize = 0;
    if (l > 0 && !L.area2) {
      const LevelGeom& P = g.lv[l - 1];
      const ResizeLut* lx = lut.data() + L.lutXOff;
      const int pw = L.w + 2 * ORBFE_EDGE;
      bool ok = true;
      std::vector<PyrWordLut> words(L.pyrWords);
      for (int k = 0; k < L.pyrWords && ok; ++k) {
        int sb[4], mn = 1 << 30;
        for (int j = 0; j < 4; ++j) {
          int x = std::min(4 * k + j, pw - 1) - ORBFE_EDGE;
          if (L.w == 1) x = 0;  // cv::borderInterpolate: a 1-px axis reflects onto itself (the loop below would not end)
          while (x < 0 || x >= L.w) x = x < 0 ? -x : 2 * (L.w - 1) - x;  // BORDER_REFLECT_101
          sb[j] = ORBFE_EDGE + lx[x].ofs;  // byte of the padded source row
          mn = std::min(mn, sb[j]);
          words[k].cpack[j] = (unsigned)(unsigned short)lx[x].c0 | ((unsigned)(unsigned short)lx[x].c1 << 16);
        }
        words[k].srcW = mn >> 2;
        words[k].sh = 8 * (mn & 3);
        words[k].sel = 0;
        for (int j = 0; j < 4; ++j) {
          const int pj = sb[j] - mn;  // offset of the pixel's left source byte inside the aligned 8-byte window
          if (pj + 1 > 7) ok = false;
          words[k].sel |= (unsigned)((pj & 7) | (((pj + 1) & 7) << 4)) << (8 * j);
        }
        if ((words[k].srcW + 2) * 4 + 3 >= P.pitch) ok = false;  // the 3-word window must stay inside the row
      }
      if (ok) {
        wlut.insert(wlut.end(), words.begin(), words.end());
        L.fastResize = 1;
        // streaming form: per 32-word strip the 16-byte-aligned start of its source bytes; box width = the widest strip
        L.pyrStrip = L.h >= ORBFE_EDGE + 1 ? 1 : 0;  // the row borders are copies of interior rows 1..19 / h-20..h-2
        L.pyrSegs = 0;  // unused: the segment count is a launch argument
        L.pyrBoxOff = (int)pyrBoxX.size();
        L.pyrBoxW = 16;
        for (int t = 0; t * 32 < L.pyrWords; ++t) {
          int lo = 1 << 30, hi = 0;
          for (int k = t * 32; k < std::min(t * 32 + 32, L.pyrWords); ++k) {
            lo = std::min(lo, words[k].srcW * 4);
            hi = std::max(hi, (words[k].srcW + 3) * 4);
          }
          const int x0 = lo & ~15;
          pyrBoxX.push_back(x0);
          L.pyrBoxW = std::max(L.pyrBoxW, align_up(hi - x0, 16));
        }
        if (L.pyrBoxW > 256) L.pyrStrip = 0;
        // per padded destination row: source rows + vertical coefficients of its REFLECT_101 image row
        L.rlutOff = (int)rlut.size();
        const ResizeLut* ly = lut.data() + L.lutYOff;
        for (int py = 0; py < L.h + 2 * ORBFE_EDGE; ++py) {
          int y = py - ORBFE_EDGE;
          if (L.h == 1) y = 0;
          while (y < 0 || y >= L.h) y = y < 0 ? -y : 2 * (L.h - 1) - y;
          PyrRowLut R;
          R.s0 = ly[y].ofs; R.s1 = std::min(ly[y].ofs + 1, P.h - 1);
          R.b0 = (unsigned)(unsigned short)ly[y].c0 << 16; R.b1 = (unsigned)(unsigned short)ly[y].c1 << 16;
          rlut.push_back(R);
        }
      }
    }
    L.pyrBlocks = (L.pyrWords * (L.h + 2 * ORBFE_EDGE) + ORBFE_PYR_THREADS - 1) / ORBFE_PYR_THREADS;
    // FAST grid (orb_extractor.cpp:714-728)
    L.maxBX = L.w - ORBFE_EDGE + 3;
    L.maxBY = L.h - ORBFE_EDGE + 3;
    const float width = (float)(L.maxBX - ORBFE_MINB), height = (float)(L.maxBY - ORBFE_MINB);
    const float W = 30;
    L.nCols = (int)(width / W);
    L.nRows = (int)(height / W);
    if (L.nCols < 1 || L.nRows < 1) {  // level too small for one cell (the reference divides by zero)
      L.nCols = L.nRows = 0; L.wCell = L.hCell = 1;
    } else {
      L.wCell = (int)std::ceil(width / L.nCols);
      L.hCell = (int)std::ceil(height / L.nRows);
    }
    L.wCellMagic = (65536 + L.wCell - 1) / L.wCell;
    L.cellBase = cellBase;
    cellBase += L.nCols * L.nRows;
    // strict 8-neighbour NMS => survivors are an independent set of the king graph
    L.cellCap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
    L.cellListOff = (unsigned)cellListOff;
    cellListOff += (size_t)L.nCols * L.nRows * L.cellCap;
    // FAST CTAs: a band segment of fG cells; the tile (<= 16 + fG*wCell + 6 px) is one 256-byte-wide TMA box.  As many cells
    // as fit: the per-CTA set-up (tile fetch, plane clearing, barriers) is paid once per tile
    L.fG = 1;
    while (L.fG < ORBFE_FAST_MAXG && (L.fG + 1) * L.wCell <= ORBFE_FAST_MAXW) ++L.fG;
    if (const char* e = getenv("ORBFE_TUNE_FAST_FG")) { const int v = atoi(e); if (v >= 1 && v <= L.fG) L.fG = v; }  // tuning only
    L.fSegs = L.nCols > 0 ? (L.nCols + L.fG - 1) / L.fG : 0;
    L.fastBase = fastBase;
    fastBase += L.fSegs * L.nRows;
    if (L.nCols > 0) {
      if (L.wCell > ORBFE_FAST_MAXW) return orbfe_fail(ORBFE_ERR_INVALID, "FAST cell too wide (%d px)", L.wCell);
      maxInnerH = std::max(maxInnerH, L.hCell);
      maxQueue = std::max(maxQueue, L.fG * L.wCell * L.hCell);
      maxCellCap = std::max(maxCellCap, L.cellCap);
      for (int i = 0; i < L.nRows; ++i)
        for (int s = 0; s < L.fSegs; ++s) fastTasks.push_back(orbfe_fast_task(l, i, s * L.fG));
    }
    // quad-tree (orb_extractor.cpp:480-531)
    L.N = ex->fpl[l];
    L.boxW = L.maxBX - ORBFE_MINB;
    L.boxH = L.maxBY - ORBFE_MINB;
    int nIni = L.boxH > 0 ? (int)std::round((float)L.boxW / (float)L.boxH) : 1;
    if (nIni < 1) nIni = 1;
    L.nIni = nIni;
    L.hX = (float)L.boxW / (float)nIni;
    L.candCap = std::max(L.nCols * L.nRows * L.cellCap, 1);
    L.candOff = (unsigned)candOff;
    candOff += (size_t)L.candCap;
    L.nodeCap = 4 * std::max(L.N, nIni) + 16;
    L.nodeOff = (unsigned)nodeOff;
    nodeOff += (size_t)L.nodeCap;
    L.outCap = std::max(L.N + 3, 4 * nIni);
    L.outOff = outOff;
    outOff += L.outCap;
    maxSort = std::max(maxSort, std::max(L.outCap, std::max(L.N, nIni)));
    L.tilesX = (L.w + ORBFE_BLUR_TW - 1) / ORBFE_BLUR_TW;  // blur: column strips of 128 px x vertical segments
    L.tilesY = std::max(1, (L.h + ORBFE_BLUR_SEG / 2) / ORBFE_BLUR_SEG);
    L.tileBase = tileBase;
    tileBase += L.tilesX * L.tilesY;
    for (int sg = 0; sg < L.tilesY; ++sg)
      for (int tx = 0; tx < L.tilesX; ++tx) blurTasks.push_back(orbfe_blur_task(l, sg, tx));
  }
  g.totalCells = cellBase;
  g.totalTiles = tileBase;
  g.totalOut = outOff;
  g.pyrStride = (unsigned)pyrOff;
  g.blurStride = (unsigned)blurOff;
  g.cellListStride = (unsigned)std::max<size_t>(cellListOff, 1);
  g.candStride = (unsigned)candOff;
  g.nodeStride = (unsigned)nodeOff;
  int sc = 1;
  while (sc < maxSort) sc <<= 1;
  g.sortCap = sc;
  g.maxNodeCap = 0;
  for (int l = 0; l < nl; ++l) g.maxNodeCap = std::max(g.maxNodeCap, g.lv[l].nodeCap);
  ex->octSmem = (size_t)sc * sizeof(unsigned long long) + 2 * (size_t)g.maxNodeCap * sizeof(int);
  // shared-memory staging of a level's candidate keys (k_octree): as many entries as keep 4 CTAs per SM resident
  ex->octStageCap = 0;
  {
#ifndef ORBFE_OCT_BUDGET_KB
#define ORBFE_OCT_BUDGET_KB 50
#endif
    const size_t budget = ORBFE_OCT_BUDGET_KB * 1024;
    if (ex->octSmem + 8 * 1024 <= budget) ex->octStageCap = (int)((budget - ex->octSmem) / 8);
    ex->octSmem += (size_t)ex->octStageCap * 8;
  }
  if (ex->octSmem > 200 * 1024)
    return orbfe_fail(ORBFE_ERR_INVALID, "nfeatures per level too large for the shared-memory sort (%d)", maxSort);
  g.totalFast = fastBase;
  ex->fastRows = maxInnerH + 6;                   // tile rows == TMA box height
  if (ex->fastRows > ORBFE_FAST_MAXROWS) return orbfe_fail(ORBFE_ERR_INVALID, "FAST cell too tall (%d rows)", ex->fastRows);
  {
    // queue capacities: what lets ORBFE_FAST_CTAS_PER_SM CTAs of this kernel share an SM (227 KB usable, 1 KB reserved and
    // ~3 KB of static shared memory per CTA) instead of the worst case (every pixel); a tile that passes more goes to the
    // kernel's dense form.  Q1 (pre-test candidates) gets 2/3, Q2 (corners) 1/3.  ORBFE_TEST_FAST_QUEUE_PCT (tests only)
    // forces small queues so that ordinary images exercise that form.
#ifndef ORBFE_FAST_CTAS_PER_SM
#define ORBFE_FAST_CTAS_PER_SM 5
#endif
    ex->fastListCap = maxCellCap;
    const size_t planes = (size_t)orbfe_fast_layout(ex->fastRows, 0, 0, ex->fastListCap).total;
    const size_t target = (227 * 1024) / ORBFE_FAST_CTAS_PER_SM - 4 * 1024;
    int cap = maxQueue;
    if (planes + 3 * (size_t)maxQueue > target) cap = (int)std::max<size_t>((target > planes ? (target - planes) / 3 : 0), (size_t)maxQueue / 8);
    if (const char* e = getenv("ORBFE_TEST_FAST_QUEUE_PCT")) { const int pc = atoi(e); if (pc >= 1 && pc <= 100) cap = std::max(maxQueue * pc / 100, 32); }
    ex->fastQ1Cap = std::min(maxQueue, cap);
    ex->fastQ2Cap = std::max(ex->fastQ1Cap / 2, 16);
  }
  ex->fastSmem = (size_t)orbfe_fast_layout(ex->fastRows, ex->fastQ1Cap, ex->fastQ2Cap, ex->fastListCap).total;
  if (ex->fastSmem > 200 * 1024) return orbfe_fail(ORBFE_ERR_INVALID, "FAST cell too large");
  ex->bestStride = (size_t)g.nodeStride * 5 / 4 + 16 * ORBFE_MAX_LEVELS;

  const size_t S = (size_t)ex->S;
  CUDA_TRY(cudaMalloc(&ex->d_img, S * g.imgStride + 16));
  CUDA_TRY(cudaMalloc(&ex->d_pyr, S * g.pyrStride));
  CUDA_TRY(cudaMalloc(&ex->d_blur, S * g.blurStride));
  CUDA_TRY(cudaMalloc(&ex->d_cellCnt, S * std::max(g.totalCells, 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->d_cellList, S * g.cellListStride * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&ex->oct.cand, S * g.candStride * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&ex->oct.knode, S * g.candStride * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->oct.cellStart, S * std::max(g.totalCells, 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->oct.nodes, S * 2 * g.nodeStride * sizeof(OctNode)));
  CUDA_TRY(cudaMalloc(&ex->oct.childCnt, S * g.nodeStride * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->oct.childSlot, S * g.nodeStride * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->oct.best, S * ex->bestStride * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&ex->oct.finSeq, S * g.totalOut * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->oct.finKey, S * g.totalOut * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->d_lvlKp, S * g.totalOut * sizeof(unsigned)));
  CUDA_TRY(cudaMalloc(&ex->d_lvlCnt, S * nl * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->d_kps, S * g.totalOut * sizeof(orbfe_kp_dev)));
  CUDA_TRY(cudaMalloc(&ex->d_desc, S * g.totalOut * 32));
  CUDA_TRY(cudaMalloc(&ex->d_nKp, S * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->d_lut, std::max<size_t>(lut.size(), 1) * sizeof(ResizeLut)));
  CUDA_TRY(cudaMalloc(&ex->d_wlut, std::max<size_t>(wlut.size(), 1) * sizeof(PyrWordLut)));
  if (!wlut.empty())
    CUDA_TRY(cudaMemcpyAsync(ex->d_wlut, wlut.data(), wlut.size() * sizeof(PyrWordLut), cudaMemcpyHostToDevice, ex->stream));
  CUDA_TRY(cudaMalloc(&ex->d_rlut, std::max<size_t>(rlut.size(), 1) * sizeof(PyrRowLut)));
  if (!rlut.empty())
    CUDA_TRY(cudaMemcpyAsync(ex->d_rlut, rlut.data(), rlut.size() * sizeof(PyrRowLut), cudaMemcpyHostToDevice, ex->stream));
  CUDA_TRY(cudaMalloc(&ex->d_err, sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->d_fastTasks, std::max<size_t>(fastTasks.size(), 1) * sizeof(unsigned)));
  if (!fastTasks.empty())
    CUDA_TRY(cudaMemcpyAsync(ex->d_fastTasks, fastTasks.data(), fastTasks.size() * sizeof(unsigned), cudaMemcpyHostToDevice, ex->stream));
  {
    // tensor maps of the padded pyramid planes, one per level, every slot as dimension 2 (orbfe_tma.cuh)
    std::vector<CUtensorMap> maps(nl);
    for (int l = 0; l < nl; ++l) {
      OrbfeTmaPlane P;
      P.base = ex->d_pyr + g.lv[l].planeOff; P.sliceStride = g.pyrStride; P.pitch = g.lv[l].pitch;
      P.rows = g.lv[l].h + 2 * ORBFE_EDGE; P.slices = (int)S; P.boxW = ORBFE_FAST_TP; P.boxH = ex->fastRows;
      const int r = orbfe_tma_encode(&maps[l], P);
      if (r != 0) return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for level %d (%d)", l, r);
    }
    CUDA_TRY(cudaMalloc(&ex->d_tmaps, nl * sizeof(CUtensorMap)));
    CUDA_TRY(cudaMemcpy(ex->d_tmaps, maps.data(), nl * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    for (int l = 0; l < nl; ++l) {
      OrbfeTmaPlane P;
      P.base = ex->d_pyr + g.lv[l].planeOff; P.sliceStride = g.pyrStride; P.pitch = g.lv[l].pitch;
      P.rows = g.lv[l].h + 2 * ORBFE_EDGE; P.slices = (int)S; P.boxW = ORBFE_BLUR_BOXW; P.boxH = ORBFE_BLUR_RB;
      const int r = orbfe_tma_encode(&maps[l], P);
      if (r != 0) return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for the blur box of level %d (%d)", l, r);
    }
    CUDA_TRY(cudaMalloc(&ex->d_tmapsBlur, nl * sizeof(CUtensorMap)));
    CUDA_TRY(cudaMemcpy(ex->d_tmapsBlur, maps.data(), nl * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    for (int l = 1; l < nl; ++l) {
      if (!g.lv[l].pyrStrip) continue;
      OrbfeTmaPlane P;
      P.base = ex->d_pyr + g.lv[l - 1].planeOff; P.sliceStride = g.pyrStride; P.pitch = g.lv[l - 1].pitch;
      P.rows = g.lv[l - 1].h + 2 * ORBFE_EDGE; P.slices = (int)S; P.boxW = g.lv[l].pyrBoxW; P.boxH = ORBFE_PYRS_RB;
      const int r = orbfe_tma_encode(&maps[l], P);
      if (r != 0) return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for the resize box of level %d (%d)", l, r);
    }
    CUDA_TRY(cudaMalloc(&ex->d_tmapsPyr, nl * sizeof(CUtensorMap)));
    CUDA_TRY(cudaMemcpy(ex->d_tmapsPyr, maps.data(), nl * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    for (int l = 0; l < nl; ++l) {
      OrbfeTmaPlane P;
      P.base = ex->d_pyr + g.lv[l].planeOff; P.sliceStride = g.pyrStride; P.pitch = g.lv[l].pitch;
      P.rows = g.lv[l].h + 2 * ORBFE_EDGE; P.slices = (int)S; P.boxW = ORBFE_DESC_PYR_BW; P.boxH = ORBFE_DESC_PYR_BH;
      const int r = orbfe_tma_encode(&maps[l], P);
      if (r != 0) return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for the moment box of level %d (%d)", l, r);
    }
    CUDA_TRY(cudaMalloc(&ex->d_tmapsDescPyr, nl * sizeof(CUtensorMap)));
    CUDA_TRY(cudaMemcpy(ex->d_tmapsDescPyr, maps.data(), nl * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    std::vector<CUtensorMap> maps2(2 * nl);  // [l]: 64-byte rows, [nl + l]: 48-byte rows (k_describe.cuh)
    for (int l = 0; l < 2 * nl; ++l) {
      const int lv = l % nl;
      OrbfeTmaPlane P;
      P.base = ex->d_blur + g.lv[lv].blurOff; P.sliceStride = g.blurStride; P.pitch = g.lv[lv].bpitch;
      P.rows = g.lv[lv].h; P.slices = (int)S; P.boxW = l < nl ? ORBFE_DESC_BLUR_BW : ORBFE_DESC_PYR_BW; P.boxH = ORBFE_DESC_BLUR_BH;
      const int r = orbfe_tma_encode(&maps2[l], P);
      if (r != 0) return orbfe_fail(ORBFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for the descriptor box of level %d (%d)", lv, r);
    }
    CUDA_TRY(cudaMalloc(&ex->d_tmapsDescBlur, 2 * nl * sizeof(CUtensorMap)));
    CUDA_TRY(cudaMemcpy(ex->d_tmapsDescBlur, maps2.data(), 2 * nl * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&ex->d_pyrBoxX, std::max<size_t>(pyrBoxX.size(), 1) * sizeof(int)));
    if (!pyrBoxX.empty())
      CUDA_TRY(cudaMemcpy(ex->d_pyrBoxX, pyrBoxX.data(), pyrBoxX.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&ex->d_blurTasks, std::max<size_t>(blurTasks.size(), 1) * sizeof(unsigned)));
    if (!blurTasks.empty())
      CUDA_TRY(cudaMemcpy(ex->d_blurTasks, blurTasks.data(), blurTasks.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
  }
  {
    // IC_Angle weights (orb_extractor.cpp:18-45, umax :393-410): byte p of the 36-byte aligned window of patch
    // row v holds pixel u = p - a - 15 (a = alignment of the row start); weight u+15 and mask 1 inside the disc
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    std::vector<uint2> icw(4 * 16 * 9);
    for (int a = 0; a < 4; ++a)
      for (int av = 0; av < 16; ++av)
        for (int i = 0; i < 9; ++i) {
          unsigned wu = 0, w1 = 0;
          for (int b = 0; b < 4; ++b) {
            const int u = 4 * i + b - a - ORBFE_HALF_PATCH;
            if (u >= -umax[av] && u <= umax[av]) { wu |= (unsigned)(u + ORBFE_HALF_PATCH) << (8 * b); w1 |= 1u << (8 * b); }
          }
          icw[(a * 16 + av) * 9 + i] = make_uint2(wu, w1);
        }
    CUDA_TRY(cudaMalloc(&ex->d_icw, icw.size() * sizeof(uint2)));
    CUDA_TRY(cudaMemcpy(ex->d_icw, icw.data(), icw.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMalloc(&ex->d_pairs, S * sizeof(StereoPair)));
  CUDA_TRY(cudaMalloc(&ex->d_uR, S * g.totalOut * sizeof(float)));
  CUDA_TRY(cudaMalloc(&ex->d_depth, S * g.totalOut * sizeof(float)));
  CUDA_TRY(cudaMalloc(&ex->d_sad, S * g.totalOut * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->d_nMatched, S * sizeof(int)));
  // stereo row table: a right keypoint spans at most 2*r+3 rows, r = 2*scale[octave] (frame.cpp:427-433)
  ex->rowCap = g.totalOut * (2 * (int)std::ceil(2.0f * ex->scale[nl - 1]) + 3);
  ex->rowSmem = (size_t)(h0 + 1) * sizeof(int);
  CUDA_TRY(cudaMalloc(&ex->d_rowStart, S * (size_t)(h0 + 1) * sizeof(int)));
  CUDA_TRY(cudaMalloc(&ex->d_rowItems, S * (size_t)ex->rowCap * sizeof(uint2)));
  CUDA_TRY(cudaMallocHost(&ex->h_img, std::min<size_t>(S, 2) * (size_t)w0 * h0));
  CUDA_TRY(cudaMallocHost(&ex->h_n, (2 * S + 3) * sizeof(int)));
  memset(ex->h_n, 0, (2 * S + 3) * sizeof(int));
  CUDA_TRY(cudaMallocHost(&ex->h_kps, S * g.totalOut * sizeof(orbfe_kp_dev)));
  CUDA_TRY(cudaMallocHost(&ex->h_desc, S * g.totalOut * 32));
  CUDA_TRY(cudaMallocHost(&ex->h_uR, S * g.totalOut * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&ex->h_depth, S * g.totalOut * sizeof(float)));
  CUDA_TRY(cudaMallocHost(&ex->h_pairs, S * sizeof(StereoPair)));
  if (!lut.empty())
    CUDA_TRY(cudaMemcpyAsync(ex->d_lut, lut.data(), lut.size() * sizeof(ResizeLut), cudaMemcpyHostToDevice, ex->stream));
  CUDA_TRY(cudaMemsetAsync(ex->d_err, 0, sizeof(int), ex->stream));
  CUDA_TRY(cudaMemsetAsync(ex->d_nKp, 0, S * sizeof(int), ex->stream));
  CUDA_TRY(cudaMemsetAsync(ex->d_nMatched, 0, S * sizeof(int), ex->stream));
  CUDA_TRY(cudaStreamSynchronize(ex->stream));  // `lut` goes out of scope
#ifndef ORBFE_EMU
  CUDA_TRY(orbfe_raise_dynamic_smem(k_fast_cells<true>, ex->device, ex->fastSmem));
  CUDA_TRY(orbfe_raise_dynamic_smem(k_fast_cells<false>, ex->device, ex->fastSmem));
  CUDA_TRY(orbfe_raise_dynamic_smem(k_octree, ex->device, ex->octSmem));
  CUDA_TRY(orbfe_raise_dynamic_smem(k_stereo_rows, ex->device, ex->rowSmem));
#endif
  ex->configured = true;
  return ORBFE_OK;
}

// stage brackets: event k of the current ring entry (0 start, 1 pyramid, 2 FAST, 3 quad-tree,
// 4 blur, 5 describe, 6 stereo search, 7 stereo median)
static int stage_event(orbfe_extractor* ex, int k) {
  if (!ex->stageTiming) return ORBFE_OK;
  if (k == 0) { ex->stageRuns++; ex->stageHas[(ex->stageRuns - 1) & 63] = 0; }
  if (ex->stageRuns == 0) return ORBFE_OK;
  const int set = (ex->stageRuns - 1) & 63;
  if (!ex->stageEv[set][k]) CUDA_TRY(cudaEventCreate(&ex->stageEv[set][k]));
  CUDA_TRY(cudaEventRecord(ex->stageEv[set][k], ex->stream));
  if (k == 5) ex->stageHas[set] = 1;
  if (k == 7 && ex->stageHas[set] == 1) ex->stageHas[set] = 2;
  return ORBFE_OK;
}

static int enqueue_extract(orbfe_extractor* ex, int n) {
  const Geom& g = ex->g;
  int rc;
  ex->slotDirty = true;
  if ((rc = stage_event(ex, 0))) return rc;
  for (int l = 0; l < g.nlevels; ++l) {
    const LevelGeom& L = g.lv[l];
    if (l == 0) {
      ORBFE_LAUNCH(ex, k_pyramid_level0, dim3(L.h + 2 * ORBFE_EDGE, n), dim3(ORBFE_PYR0_THREADS), 0, g, ex->d_img, ex->d_pyr);
      continue;
    }
    if (L.pyrStrip && !getenv("ORBFE_TUNE_PYR_OLD")) {
      // vertical segments: about ORBFE_PYRS_SEG rows each, shorter on the small levels so that a launch still brings a few
      // thousand warps (a warp marches its rows one after the other: the launch cannot end before the longest march)
      const int strips = (L.pyrWords + 31) / 32;
      int segT = ORBFE_PYRS_SEG, minSeg = 16, wantWarps = 0;  // A/B on B200 (128 frames): forcing 3000 / 6000 / 12000 warps per launch -> 0.253 / 0.262 / 0.291 ms (0: 0.251)
      if (const char* e = getenv("ORBFE_TUNE_PYR_SEG")) { const int v = atoi(e); if (v >= 8 && v <= 512) segT = v; }      // tuning only
      if (const char* e = getenv("ORBFE_TUNE_PYR_WARPS")) { const int v = atoi(e); if (v >= 0) wantWarps = v; }           // tuning only
      int segs = (L.h + segT - 1) / segT;
      segs = std::max(segs, std::min((L.h + minSeg - 1) / minSeg, (wantWarps + strips * n - 1) / (strips * n)));
      // latency path (a frame or a pair per launch): 56-row marches leave most SMs idle and the launch lasts as long as one
      // march; shorter segments (>= 12 rows) until the launch brings about 4 warps per SM
      if (!getenv("ORBFE_TUNE_PYR_SEG") && !getenv("ORBFE_TUNE_PYR_WARPS"))
        while ((long long)strips * segs * n < 600 && (L.h + segs) / (segs + 1) >= 12) ++segs;
      const int segH = (L.h + segs - 1) / segs;
      const size_t warpBytes = ((size_t)ORBFE_PYRS_RING * L.pyrBoxW + (size_t)segH * sizeof(PyrRowLut) + 127) & ~(size_t)127;
      const size_t smem = warpBytes * ORBFE_PYRS_WPC;
      const int tasks = strips * segs;
      ORBFE_LAUNCH(ex, k_pyramid_strip, dim3((tasks + ORBFE_PYRS_WPC - 1) / ORBFE_PYRS_WPC, n), dim3(ORBFE_PYRS_THREADS), smem, g, l, ex->d_pyr,
                   ex->d_tmapsPyr, ex->d_rlut, ex->d_wlut, ex->d_pyrBoxX, segs);
    } else if (L.fastResize) {
      // a strip is a chain of dependent loads: long strips (row re-use) only when the launch still fills
      // the GPU with warps several times over; otherwise short strips, more warps, less latency
      const int strips = (L.pyrWords + 31) / 32, ph = L.h + 2 * ORBFE_EDGE;
      int stripRows = ORBFE_PYR_ROWS;
      if (const char* e = getenv("ORBFE_TUNE_PYR_ROWS")) { const int v = atoi(e); if (v >= 2 && v <= 64) stripRows = v; }  // tuning only; A/B at 64 pairs: 4 -> 0.274 ms, 6 -> 0.265, 8 -> 0.255, 16 -> 0.254
      while (stripRows > 2 && (long long)strips * ((ph + stripRows - 1) / stripRows) * n < 1LL * 148 * 64) stripRows >>= 1;
      const int tasks = ((L.pyrWords + 31) / 32) * ((L.h + 2 * ORBFE_EDGE + stripRows - 1) / stripRows);
      ORBFE_LAUNCH(ex, k_pyramid_resize, dim3((tasks + ORBFE_PYR_THREADS / 32 - 1) / (ORBFE_PYR_THREADS / 32), n),
                   dim3(ORBFE_PYR_THREADS), 0, g, l, stripRows, ex->d_pyr, ex->d_rlut, ex->d_wlut);
    } else {
      ORBFE_LAUNCH(ex, k_pyramid_level, dim3(L.pyrBlocks, n), dim3(ORBFE_PYR_THREADS), 0, g, l, ex->d_img, ex->d_pyr,
                   ex->d_lut);
    }
  }
  if ((rc = stage_event(ex, 1))) return rc;
  if (g.totalFast > 0) {
    auto kfast = (g.iniTh < 128 && g.minTh < 128) ? k_fast_cells<true> : k_fast_cells<false>;
    ORBFE_LAUNCH(ex, kfast, dim3(g.totalFast, n), dim3(ORBFE_FAST_THREADS), ex->fastSmem, g, ex->d_pyr, ex->d_tmaps,
                 ex->d_fastTasks, ex->d_cellCnt, ex->d_cellList, ex->fastRows, ex->fastQ1Cap, ex->fastQ2Cap, ex->fastListCap);
  }
  if ((rc = stage_event(ex, 2))) return rc;
  ORBFE_LAUNCH(ex, k_octree, dim3(n, g.nlevels), dim3(ORBFE_OCT_THREADS), ex->octSmem, g, ex->d_cellCnt, ex->d_cellList,
               ex->oct, ex->d_lvlKp, ex->d_lvlCnt, ex->d_err, ex->octStageCap);
  if ((rc = stage_event(ex, 3))) return rc;
  ORBFE_LAUNCH(ex, k_blur, dim3(g.totalTiles, n), dim3(ORBFE_BLUR_THREADS), 0, g, ex->d_pyr, ex->d_tmapsBlur, ex->d_blurTasks,
               ex->d_blur, orbfe_blur_consts());  // one CTA (one warp) per 128-px column strip of a level
  if ((rc = stage_event(ex, 4))) return rc;
  {
    int kpw = n >= 4 ? 8 : 4;  // keypoints per warp (see k_orient_describe).  A/B at 128 frames, TMA form: 4 -> 0.203 ms, 6 -> 0.194, 8 -> 0.191,
                               // 12 -> 0.190, 16 -> 0.205, 32 -> 0.215
    if (const char* e = getenv("ORBFE_TUNE_KPW")) { const int v = atoi(e); if (v >= 1 && v <= 32) kpw = v; }  // tuning only
    const int warps = (g.totalOut + kpw - 1) / kpw, wpc = ORBFE_DESC_THREADS / 32;
    ORBFE_LAUNCH(ex, k_orient_describe, dim3((warps + wpc - 1) / wpc, n), dim3(ORBFE_DESC_THREADS), 0, g, ex->d_pyr, ex->d_blur,
                 ex->d_lvlKp, ex->d_lvlCnt, ex->d_kps, ex->d_desc, ex->d_nKp, kpw, ex->d_icw, ex->d_tmapsDescPyr, ex->d_tmapsDescBlur);
  }
  if ((rc = stage_event(ex, 5))) return rc;
  CUDA_TRY(cudaGetLastError());
  return ORBFE_OK;
}

static int check_images(orbfe_extractor* ex, int first, int n) {
  if (!ex) return orbfe_fail(ORBFE_ERR_INVALID, "null extractor handle");
  if (n < 0 || first < 0 || first + n > ex->S)
    return orbfe_fail(ORBFE_ERR_INVALID, "slots [%d,%d) exceed max_images=%d", first, first + n, ex->S);
  return ORBFE_OK;
}

int orbfe_internal_slot_view(orbfe_extractor* ex, int slot, OrbfeSlotView* out) {
  if (!ex || !out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  if (!ex->configured || slot < 0 || slot >= ex->S) return orbfe_fail(ORBFE_ERR_INVALID, "slot %d has no extraction results", slot);
  const Geom& g = ex->g;
  out->device = ex->device; out->stream = ex->stream;
  out->kps = ex->d_kps + (size_t)slot * g.totalOut;
  out->desc = ex->d_desc + (size_t)slot * g.totalOut * 32;
  out->uR = ex->d_uR ? ex->d_uR + (size_t)slot * g.totalOut : nullptr;
  out->nKp = ex->d_nKp + slot;
  out->capacity = g.totalOut; out->nlevels = g.nlevels; out->w = g.w0; out->h = g.h0;
  for (int l = 0; l < g.nlevels && l < 16; ++l) out->scale[l] = ex->scale[l];
  return ORBFE_OK;
}

extern "C" {

const char* orbfe_last_error(void) { return t_last_error.c_str(); }
const char* orbfe_version(void) {
#ifdef ORBFE_EMU
  return "orbfe 0.1 (EMULATED TEST BUILD - not a product library)";
#else
  return "orbfe 0.1 (sm_100a)";
#endif
}
int orbfe_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int orbfe_extractor_create(const orbfe_params* p, int device, orbfe_extractor** out) {
  if (!p || !out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (p->nlevels < 1 || p->nlevels > ORBFE_MAX_LEVELS)
    return orbfe_fail(ORBFE_ERR_INVALID, "nlevels must be in [1,%d]", ORBFE_MAX_LEVELS);
  if (p->nfeatures < 1 || !(p->scale_factor > 1.0f) || p->ini_th_fast < 1 || p->min_th_fast < 1 ||
      p->ini_th_fast > 255 || p->min_th_fast > p->ini_th_fast)
    return orbfe_fail(ORBFE_ERR_INVALID, "bad ORB parameters");
  const int ndev = orbfe_device_count();
  if (device < 0 || device >= ndev)
    return orbfe_fail(ORBFE_ERR_CUDA, "CUDA device %d not available (%d visible); this library has no CPU path", device, ndev);
  orbfe_extractor* ex = new (std::nothrow) orbfe_extractor();
  if (!ex) return orbfe_fail(ORBFE_ERR_NOMEM, "out of host memory");
  ex->device = device;
  ex->p = *p;
  ex->S = p->max_images > 0 ? p->max_images : 1;
  // scale tables and per-level feature quota (orb_extractor.cpp:356-387); scaleFactor is a double
  // member holding the float argument (orb_extractor.h:79)
  const double sf = (double)p->scale_factor;
  const int nl = p->nlevels;
  ex->scale[0] = 1.0f; ex->sigma2[0] = 1.0f;
  for (int i = 1; i < nl; ++i) {
    ex->scale[i] = (float)((double)ex->scale[i - 1] * sf);
    ex->sigma2[i] = ex->scale[i] * ex->scale[i];
  }
  for (int i = 0; i < nl; ++i) { ex->invScale[i] = 1.0f / ex->scale[i]; ex->invSigma2[i] = 1.0f / ex->sigma2[i]; }
  const float factor = (float)(1.0 / sf);
  float nDesired = p->nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl));
  int sum = 0;
  for (int l = 0; l < nl - 1; ++l) {
    ex->fpl[l] = cv_round_f(nDesired);
    sum += ex->fpl[l];
    nDesired *= factor;
  }
  ex->fpl[nl - 1] = std::max(p->nfeatures - sum, 0);
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking);
  for (int i = 0; i < 64 && e == cudaSuccess; ++i) e = cudaEventCreate(&ex->ev[i]);
  if (e != cudaSuccess) {
    delete ex;
    return orbfe_fail(ORBFE_ERR_CUDA, "device %d setup failed: %s", device, cudaGetErrorString(e));
  }
  if (p->max_width > 0 && p->max_height > 0) {
    const int rc = configure(ex, p->max_width, p->max_height);
    if (rc != ORBFE_OK) { orbfe_extractor_destroy(ex); return rc; }
  }
  *out = ex;
  return ORBFE_OK;
}

int orbfe_extractor_destroy(orbfe_extractor* ex) {
  if (!ex) return ORBFE_OK;
  cudaSetDevice(ex->device);
  if (ex->stream) cudaStreamSynchronize(ex->stream);
  free_arena(ex);
  for (int i = 0; i < 64; ++i) if (ex->ev[i]) cudaEventDestroy(ex->ev[i]);
  for (int i = 0; i < 64; ++i)
    for (int k = 0; k < 8; ++k) if (ex->stageEv[i][k]) cudaEventDestroy(ex->stageEv[i][k]);
  if (ex->stream) cudaStreamDestroy(ex->stream);
  delete ex;
  return ORBFE_OK;
}

int orbfe_extractor_tables(const orbfe_extractor* ex, int* nlevels, float* scale, float* inv_scale, float* sigma2,
                           float* inv_sigma2, int32_t* features_per_level) {
  if (!ex) return orbfe_fail(ORBFE_ERR_INVALID, "null extractor handle");
  if (nlevels) *nlevels = ex->p.nlevels;
  for (int i = 0; i < ex->p.nlevels; ++i) {
    if (scale) scale[i] = ex->scale[i];
    if (inv_scale) inv_scale[i] = ex->invScale[i];
    if (sigma2) sigma2[i] = ex->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = ex->invSigma2[i];
    if (features_per_level) features_per_level[i] = ex->fpl[i];
  }
  return ORBFE_OK;
}

int orbfe_extractor_max_keypoints(const orbfe_extractor* ex) {
  if (!ex) return orbfe_fail(ORBFE_ERR_INVALID, "null extractor handle");
  if (ex->configured) return ex->g.totalOut;
  // before the first image the aspect ratio (nIni) is unknown: N_l + 3 per level, and 4*nIni for
  // levels whose quota is tiny; 64 per level covers aspect ratios up to 16:1
  return ex->p.nfeatures + 64 * ex->p.nlevels;
}

int orbfe_upload(orbfe_extractor* ex, int first_slot, const uint8_t* const* imgs, int n_imgs, int w, int h,
                 size_t stride) {
  int rc = check_images(ex, first_slot, n_imgs);
  if (rc) return rc;
  if (!imgs || w <= 0 || h <= 0 || stride < (size_t)w) return orbfe_fail(ORBFE_ERR_INVALID, "bad image arguments");
  CUDA_TRY(cudaSetDevice(ex->device));
  if ((rc = configure(ex, w, h))) return rc;
  for (int i = 0; i < n_imgs; ++i)
    if (!imgs[i]) return orbfe_fail(ORBFE_ERR_INVALID, "null image %d", i);
  const size_t bytes = (size_t)w * h;
  if (stride == (size_t)w) {
    // tightly packed frames: one 1-D copy per run of frames that are contiguous on the host too
    int i = 0;
    while (i < n_imgs) {
      int j = i + 1;
      if (ex->g.imgStride == bytes)
        while (j < n_imgs && imgs[j] == imgs[j - 1] + bytes) ++j;
      CUDA_TRY(cudaMemcpyAsync(ex->d_img + (size_t)(first_slot + i) * ex->g.imgStride, imgs[i], (size_t)(j - i) * bytes,
                               cudaMemcpyHostToDevice, ex->stream));
      i = j;
    }
  } else {
    for (int i = 0; i < n_imgs; ++i)
      CUDA_TRY(cudaMemcpy2DAsync(ex->d_img + (size_t)(first_slot + i) * ex->g.imgStride, ex->g.imgPitch, imgs[i], stride,
                                 (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ex->stream));
  }
  return ORBFE_OK;
}

int orbfe_upload_color(orbfe_extractor* ex, int first_slot, const uint8_t* const* imgs, int n_imgs, int w, int h,
                       size_t stride, int channels, int rgb_order) {
  int rc = check_images(ex, first_slot, n_imgs);
  if (rc) return rc;
  if (!imgs || w <= 0 || h <= 0 || (channels != 3 && channels != 4) || stride < (size_t)w * channels)
    return orbfe_fail(ORBFE_ERR_INVALID, "bad colour image arguments");
  CUDA_TRY(cudaSetDevice(ex->device));
  if ((rc = configure(ex, w, h))) return rc;
  const size_t rowBytes = (size_t)w * channels, bytes = rowBytes * h;
  const size_t cstride = align_up_sz(bytes, 16) + 16;
  if (!ex->d_color || ex->colorStride < cstride) {
    CUDA_TRY(cudaStreamSynchronize(ex->stream));
    cudaFree(ex->d_color);
    ex->d_color = nullptr;
    CUDA_TRY(cudaMalloc(&ex->d_color, (size_t)ex->S * cstride));
    ex->colorStride = cstride;
  }
  for (int i = 0; i < n_imgs; ++i) {
    if (!imgs[i]) return orbfe_fail(ORBFE_ERR_INVALID, "null image %d", i);
    uint8_t* d = ex->d_color + (size_t)(first_slot + i) * ex->colorStride;
    if (stride == rowBytes) CUDA_TRY(cudaMemcpyAsync(d, imgs[i], bytes, cudaMemcpyHostToDevice, ex->stream));
    else CUDA_TRY(cudaMemcpy2DAsync(d, rowBytes, imgs[i], stride, rowBytes, (size_t)h, cudaMemcpyHostToDevice, ex->stream));
  }
  const int nWords = (w * h + 3) / 4;
  ORBFE_LAUNCH(ex, k_gray, dim3((nWords + ORBFE_GRAY_THREADS - 1) / ORBFE_GRAY_THREADS, n_imgs), dim3(ORBFE_GRAY_THREADS), 0,
               ex->d_color + (size_t)first_slot * ex->colorStride, ex->colorStride, ex->d_img + (size_t)first_slot * ex->g.imgStride,
               (size_t)ex->g.imgStride, nWords, channels, rgb_order ? 1 : 0);
  CUDA_TRY(cudaGetLastError());
  return ORBFE_OK;
}

int orbfe_run(orbfe_extractor* ex, int n_imgs) {
  int rc = check_images(ex, 0, n_imgs);
  if (rc) return rc;
  if (!ex->configured) return orbfe_fail(ORBFE_ERR_INVALID, "orbfe_run before any orbfe_upload");
  if (n_imgs == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(ex->device));
  return enqueue_extract(ex, n_imgs);
}

// fills the StereoPair table for pairs (slot 2p, slot 2p+1) of ONE handle
static void fill_pairs_same_handle(orbfe_extractor* ex, int n_pairs) {
  const Geom& g = ex->g;
  for (int p = 0; p < n_pairs; ++p) {
    const size_t a = 2 * (size_t)p, b = a + 1;
    StereoPair& P = ex->h_pairs[p];
    P.pyrL = ex->d_pyr + a * g.pyrStride; P.pyrR = ex->d_pyr + b * g.pyrStride;
    P.kpL = ex->d_kps + a * g.totalOut; P.kpR = ex->d_kps + b * g.totalOut;
    P.descL = ex->d_desc + a * g.totalOut * 32; P.descR = ex->d_desc + b * g.totalOut * 32;
    P.nL = ex->d_nKp + a; P.nR = ex->d_nKp + b;
    P.uR = ex->d_uR + a * g.totalOut; P.depth = ex->d_depth + a * g.totalOut; P.sad = ex->d_sad + a * g.totalOut;
    P.rowStart = ex->d_rowStart + (size_t)p * (g.h0 + 1); P.rowItems = ex->d_rowItems + (size_t)p * ex->rowCap;
    P.rowCap = ex->rowCap;
  }
}

static int enqueue_stereo(orbfe_extractor* ex, int n_pairs, float bf, float baseline) {
  const Geom& g = ex->g;
  // one CTA per pair; a lone pair (the latency path) gets 1024 threads: its two passes over the right keypoints are latency-bound
  ORBFE_LAUNCH(ex, k_stereo_rows, dim3(n_pairs), dim3(n_pairs <= 8 ? 1024 : 256), ex->rowSmem, g, ex->d_pairs, g.totalOut);
  ORBFE_LAUNCH(ex, k_stereo_search, dim3((g.totalOut + ORBFE_ST_THREADS / 32 - 1) / (ORBFE_ST_THREADS / 32), n_pairs),
               dim3(ORBFE_ST_THREADS), 0, g, ex->d_pairs, bf, baseline, g.totalOut);
  int rc;
  if ((rc = stage_event(ex, 6))) return rc;
  ORBFE_LAUNCH(ex, k_stereo_median, dim3(n_pairs), dim3(ORBFE_ST_THREADS), 0, ex->d_pairs, g.totalOut, ex->d_nMatched);
  if ((rc = stage_event(ex, 7))) return rc;
  CUDA_TRY(cudaGetLastError());
  return ORBFE_OK;
}

int orbfe_run_stereo(orbfe_extractor* ex, int n_pairs, float bf, float baseline) {
  int rc = check_images(ex, 0, 2 * n_pairs);
  if (rc) return rc;
  if (!ex->configured) return orbfe_fail(ORBFE_ERR_INVALID, "orbfe_run_stereo before any orbfe_upload");
  if (!(baseline > 0.f) || !(bf > 0.f)) return orbfe_fail(ORBFE_ERR_INVALID, "bf and baseline must be positive");
  if (n_pairs == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(ex->device));
  if (ex->pairsCached < n_pairs) {
    // h_pairs is rewritten only when the previous table has been consumed
    CUDA_TRY(cudaStreamSynchronize(ex->stream));
    fill_pairs_same_handle(ex, n_pairs);
    CUDA_TRY(cudaMemcpyAsync(ex->d_pairs, ex->h_pairs, (size_t)n_pairs * sizeof(StereoPair), cudaMemcpyHostToDevice, ex->stream));
    ex->pairsCached = n_pairs;
  }
  return enqueue_stereo(ex, n_pairs, bf, baseline);
}

// D2H of the results of n images into the handle's pinned staging (enqueue only)
static int enqueue_download(orbfe_extractor* ex, size_t n, bool kps, bool desc, bool u_right, bool depth) {
  const size_t S = (size_t)ex->S, T = (size_t)ex->g.totalOut;
  CUDA_TRY(cudaMemcpyAsync(ex->h_n, ex->d_nKp, n * sizeof(int), cudaMemcpyDeviceToHost, ex->stream));
  CUDA_TRY(cudaMemcpyAsync(ex->h_n + 2 * S, ex->d_err, sizeof(int), cudaMemcpyDeviceToHost, ex->stream));
  if (kps) CUDA_TRY(cudaMemcpyAsync(ex->h_kps, ex->d_kps, n * T * sizeof(orbfe_kp_dev), cudaMemcpyDeviceToHost, ex->stream));
  if (desc) CUDA_TRY(cudaMemcpyAsync(ex->h_desc, ex->d_desc, n * T * 32, cudaMemcpyDeviceToHost, ex->stream));
  if (u_right) CUDA_TRY(cudaMemcpyAsync(ex->h_uR, ex->d_uR, n * T * sizeof(float), cudaMemcpyDeviceToHost, ex->stream));
  if (depth) CUDA_TRY(cudaMemcpyAsync(ex->h_depth, ex->d_depth, n * T * sizeof(float), cudaMemcpyDeviceToHost, ex->stream));
  return ORBFE_OK;
}

// after the stream has been synchronised: pinned staging -> the caller's arrays
static int deliver_download(orbfe_extractor* ex, size_t n, orbfe_keypoint* kps, uint8_t* desc, int capacity, int* n_out,
                            float* u_right, float* depth) {
  const size_t S = (size_t)ex->S, T = (size_t)ex->g.totalOut;
  if (ex->h_n[2 * S] != 0) {
    const int flag = ex->h_n[2 * S];
    ex->h_n[2 * S] = 0;
    cudaMemsetAsync(ex->d_err, 0, sizeof(int), ex->stream);
    return orbfe_fail(ORBFE_ERR_CUDA, "quad-tree kernel reported an internal capacity error (flag %d)", flag);
  }
  int status = ORBFE_OK;
  for (size_t i = 0; i < n; ++i) {
    const int c = ex->h_n[i];
    n_out[i] = c;
    if (c > capacity) { status = ORBFE_ERR_CAPACITY; continue; }
    if (kps) memcpy(kps + i * (size_t)capacity, ex->h_kps + i * T, (size_t)c * sizeof(orbfe_keypoint));
    if (desc) memcpy(desc + i * (size_t)capacity * 32, ex->h_desc + i * T * 32, (size_t)c * 32);
    if ((i & 1) == 0) {
      if (u_right) memcpy(u_right + i * (size_t)capacity, ex->h_uR + i * T, (size_t)c * sizeof(float));
      if (depth) memcpy(depth + i * (size_t)capacity, ex->h_depth + i * T, (size_t)c * sizeof(float));
    }
  }
  if (status != ORBFE_OK) return orbfe_fail(status, "capacity %d too small for the keypoint count", capacity);
  if (kps && desc) ex->slotDirty = false;   // the pinned copies now mirror the device rows (orbfe_stereo_match)
  return ORBFE_OK;
}

int orbfe_download(orbfe_extractor* ex, int n_imgs, orbfe_keypoint* kps, uint8_t* desc, int capacity, int* n_out,
                   float* u_right, float* depth) {
  int rc = check_images(ex, 0, n_imgs);
  if (rc) return rc;
  if (!ex->configured) return orbfe_fail(ORBFE_ERR_INVALID, "orbfe_download before any orbfe_upload");
  if (!n_out || capacity < 0) return orbfe_fail(ORBFE_ERR_INVALID, "bad output arguments");
  if (n_imgs == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(ex->device));
  if ((rc = enqueue_download(ex, (size_t)n_imgs, kps != nullptr, desc != nullptr, u_right != nullptr, depth != nullptr))) return rc;
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  return deliver_download(ex, (size_t)n_imgs, kps, desc, capacity, n_out, u_right, depth);
}

int orbfe_download_async(orbfe_extractor* ex, int n_imgs, orbfe_keypoint* kps, uint8_t* desc, int capacity, int* n_out,
                         float* u_right, float* depth) {
  int rc = check_images(ex, 0, n_imgs);
  if (rc) return rc;
  if (!ex->configured) return orbfe_fail(ORBFE_ERR_INVALID, "orbfe_download_async before any orbfe_upload");
  if (!n_out) return orbfe_fail(ORBFE_ERR_INVALID, "bad output arguments");
  if (capacity != ex->g.totalOut)
    return orbfe_fail(ORBFE_ERR_INVALID, "orbfe_download_async needs capacity == orbfe_extractor_max_keypoints() (%d)", ex->g.totalOut);
  if (n_imgs == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(ex->device));
  const size_t T = (size_t)ex->g.totalOut, n = (size_t)n_imgs;
  CUDA_TRY(cudaMemcpyAsync(n_out, ex->d_nKp, n * sizeof(int), cudaMemcpyDeviceToHost, ex->stream));
  CUDA_TRY(cudaMemcpyAsync(ex->h_n + 2 * (size_t)ex->S, ex->d_err, sizeof(int), cudaMemcpyDeviceToHost, ex->stream));  // checked by orbfe_sync
  if (kps) CUDA_TRY(cudaMemcpyAsync(kps, ex->d_kps, n * T * sizeof(orbfe_kp_dev), cudaMemcpyDeviceToHost, ex->stream));
  if (desc) CUDA_TRY(cudaMemcpyAsync(desc, ex->d_desc, n * T * 32, cudaMemcpyDeviceToHost, ex->stream));
  if (u_right) CUDA_TRY(cudaMemcpyAsync(u_right, ex->d_uR, n * T * sizeof(float), cudaMemcpyDeviceToHost, ex->stream));
  if (depth) CUDA_TRY(cudaMemcpyAsync(depth, ex->d_depth, n * T * sizeof(float), cudaMemcpyDeviceToHost, ex->stream));
  return ORBFE_OK;
}

int orbfe_pinned_alloc(size_t bytes, void** out) {
  if (!out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *out = nullptr;
  CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return ORBFE_OK;
}

int orbfe_pinned_free(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return ORBFE_OK;
}

int orbfe_sync(orbfe_extractor* ex) {
  if (!ex) return orbfe_fail(ORBFE_ERR_INVALID, "null extractor handle");
  CUDA_TRY(cudaSetDevice(ex->device));
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  if (ex->configured && ex->h_n[2 * (size_t)ex->S] != 0) {
    const int flag = ex->h_n[2 * (size_t)ex->S];
    ex->h_n[2 * (size_t)ex->S] = 0;
    cudaMemsetAsync(ex->d_err, 0, sizeof(int), ex->stream);
    return orbfe_fail(ORBFE_ERR_CUDA, "quad-tree kernel reported an internal capacity error (flag %d)", flag);
  }
  return ORBFE_OK;
}

#ifndef ORBFE_EMU
// The latency path: <= 2 frames per call.  The first call of a geometry captures  H2D (from pinned staging) -> pyramid ->
// FAST -> quad-tree -> blur -> orientation + rBRIEF -> D2H (into pinned staging)  on the handle's stream into a CUDA graph;
// every later call is one host copy of the frame(s) into the staging buffer, ONE cudaGraphLaunch and one synchronisation
// instead of 1 + 15 + 4 stream operations.  Capture is thread-local: the other extraction thread of a stereo Frame
// (frame.cpp:86-89) keeps issuing its own work meanwhile.
static int extract_graph(orbfe_extractor* ex, const uint8_t* const* imgs, int n, int w, int h, size_t stride, orbfe_keypoint* kps,
                         uint8_t* desc, int capacity, int* n_out, bool* used) {
  *used = false;
  if (n < 1 || n > 2 || n > ex->S || ex->stageTiming || getenv("ORBFE_NO_GRAPH")) return ORBFE_OK;
  int rc = check_images(ex, 0, n);
  if (rc) return rc;
  if (!imgs || w <= 0 || h <= 0 || stride < (size_t)w || !n_out || capacity < 0) return ORBFE_OK;  // the plain path reports it
  for (int i = 0; i < n; ++i) if (!imgs[i]) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(ex->device));
  if ((rc = configure(ex, w, h))) return rc;
  const size_t bytes = (size_t)w * h;
  const int key = (kps ? 1 : 0) | (desc ? 2 : 0);
  cudaGraphExec_t& exec = ex->xGraph[n - 1][key];
  if (!exec) {
    CUDA_TRY(cudaStreamSynchronize(ex->stream));
    const long long l0 = ex->launches;
    CUDA_TRY(cudaStreamBeginCapture(ex->stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n && e == cudaSuccess; ++i)
      e = cudaMemcpyAsync(ex->d_img + (size_t)i * ex->g.imgStride, ex->h_img + (size_t)i * bytes, bytes, cudaMemcpyHostToDevice, ex->stream);
    rc = e == cudaSuccess ? enqueue_extract(ex, n) : ORBFE_ERR_CUDA;
    if (rc == ORBFE_OK) rc = enqueue_download(ex, (size_t)n, kps != nullptr, desc != nullptr, false, false);
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(ex->stream, &graph);
    ex->xGraphLaunches[n - 1][key] = (int)(ex->launches - l0);
    ex->launches = l0;
    if (rc == ORBFE_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (rc != ORBFE_OK || e != cudaSuccess || !exec) {  // not capturable here: the plain path does the work
      exec = nullptr;
      cudaGetLastError();
      return ORBFE_OK;
    }
  }
  for (int i = 0; i < n; ++i) {
    uint8_t* d = ex->h_img + (size_t)i * bytes;
    if (stride == (size_t)w) memcpy(d, imgs[i], bytes);
    else for (int y = 0; y < h; ++y) memcpy(d + (size_t)y * w, imgs[i] + (size_t)y * stride, (size_t)w);
  }
  ex->slotDirty = true;
  CUDA_TRY(cudaGraphLaunch(exec, ex->stream));
  ex->launches += ex->xGraphLaunches[n - 1][key];
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  *used = true;
  return deliver_download(ex, (size_t)n, kps, desc, capacity, n_out, nullptr, nullptr);
}
#endif

int orbfe_extract_batch(orbfe_extractor* ex, const uint8_t* const* imgs, int n_imgs, int w, int h, size_t stride,
                        orbfe_keypoint* kps, uint8_t* desc, int capacity, int* n_out) {
  int rc;
#ifndef ORBFE_EMU
  if (ex && n_imgs >= 1 && n_imgs <= 2) {
    bool used = false;
    if ((rc = extract_graph(ex, imgs, n_imgs, w, h, stride, kps, desc, capacity, n_out, &used)) || used) return rc;
  }
#endif
  if ((rc = orbfe_upload(ex, 0, imgs, n_imgs, w, h, stride))) return rc;
  if ((rc = orbfe_run(ex, n_imgs))) return rc;
  return orbfe_download(ex, n_imgs, kps, desc, capacity, n_out, nullptr, nullptr);
}

int orbfe_extract(orbfe_extractor* ex, const uint8_t* img, int w, int h, size_t stride, orbfe_keypoint* kps,
                  uint8_t* desc, int capacity, int* n_out) {
  if (!ex || !n_out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *n_out = 0;
  if (!img || w <= 0 || h <= 0) return ORBFE_OK;  // _image.empty() => return (orb_extractor.cpp:990-991)
  const uint8_t* one[1] = {img};
  return orbfe_extract_batch(ex, one, 1, w, h, stride, kps, desc, capacity, n_out);
}

int orbfe_pyramid_level(orbfe_extractor* ex, int slot, int level, uint8_t* dst, size_t dst_stride, int* w, int* h) {
  int rc = check_images(ex, slot, 1);
  if (rc) return rc;
  if (!ex->configured) return orbfe_fail(ORBFE_ERR_INVALID, "no image has been processed yet");
  if (level < 0 || level >= ex->g.nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "bad level %d", level);
  const LevelGeom& L = ex->g.lv[level];
  if (w) *w = L.w;
  if (h) *h = L.h;
  if (!dst) return ORBFE_OK;
  if (dst_stride < (size_t)L.w) return orbfe_fail(ORBFE_ERR_INVALID, "dst_stride too small");
  CUDA_TRY(cudaSetDevice(ex->device));
  const uint8_t* src = ex->d_pyr + (size_t)slot * ex->g.pyrStride + L.planeOff + (size_t)ORBFE_EDGE * L.pitch + ORBFE_EDGE;
  CUDA_TRY(cudaMemcpy2DAsync(dst, dst_stride, src, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, ex->stream));
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  return ORBFE_OK;
}

int orbfe_event_record(orbfe_extractor* ex, int slot) {
  if (!ex || slot < 0 || slot >= 64) return orbfe_fail(ORBFE_ERR_INVALID, "bad event slot");
  CUDA_TRY(cudaSetDevice(ex->device));
  CUDA_TRY(cudaEventRecord(ex->ev[slot], ex->stream));
  return ORBFE_OK;
}
int orbfe_event_elapsed_ms(orbfe_extractor* ex, int a, int b, float* ms) {
  if (!ex || !ms || a < 0 || a >= 64 || b < 0 || b >= 64) return orbfe_fail(ORBFE_ERR_INVALID, "bad event slot");
  CUDA_TRY(cudaSetDevice(ex->device));
  CUDA_TRY(cudaEventElapsedTime(ms, ex->ev[a], ex->ev[b]));
  return ORBFE_OK;
}
int orbfe_set_stage_timing(orbfe_extractor* ex, int enabled) {
  if (!ex) return orbfe_fail(ORBFE_ERR_INVALID, "null extractor handle");
  ex->stageTiming = enabled != 0;
  return ORBFE_OK;
}
int orbfe_stage_summary(orbfe_extractor* ex, float* ms_sum, int* n_runs) {
  if (!ex || !ms_sum || !n_runs) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(ex->device));
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  for (int k = 0; k < ORBFE_NUM_STAGES; ++k) ms_sum[k] = 0.f;
  const int runs = ex->stageRuns < 64 ? ex->stageRuns : 64;
  for (int r = 0; r < runs; ++r) {
    const int set = (ex->stageRuns - 1 - r) & 63;
    const int last = ex->stageHas[set] == 2 ? 7 : (ex->stageHas[set] == 1 ? 5 : 0);
    for (int k = 1; k <= last; ++k) {
      if (k == 6) {  // the stereo launch follows the describe bracket (+ the pair-table copy)
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ex->stageEv[set][5], ex->stageEv[set][6]));
        ms_sum[5] += ms;
        continue;
      }
      float ms = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&ms, ex->stageEv[set][k - 1], ex->stageEv[set][k]));
      ms_sum[k - 1] += ms;
    }
  }
  *n_runs = runs;
  ex->stageRuns = 0;
  return ORBFE_OK;
}
long long orbfe_launch_count(const orbfe_extractor* ex) { return ex ? ex->launches : 0; }

// ---- debug taps (per-stage parity tests) ------------------------------------------------------
static void unpack_kp(unsigned pk, orbfe_keypoint* k) {
  k->x = (float)ORBFE_PX(pk); k->y = (float)ORBFE_PY(pk); k->size = 7.f; k->angle = -1.f;
  k->response = (float)ORBFE_PS(pk); k->octave = 0; k->class_id = -1;
}

int orbfe_debug_candidates(orbfe_extractor* ex, int slot, int level, orbfe_keypoint* out, int capacity, int* n_out) {
  int rc = check_images(ex, slot, 1);
  if (rc) return rc;
  if (!ex->configured || level < 0 || level >= ex->g.nlevels || !n_out) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  CUDA_TRY(cudaSetDevice(ex->device));
  const Geom& g = ex->g;
  const LevelGeom& L = g.lv[level];
  const int nCells = L.nCols * L.nRows;
  std::vector<int> cnt(std::max(nCells, 1));
  std::vector<unsigned> list((size_t)std::max(nCells, 1) * L.cellCap);
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  if (nCells > 0) {
    CUDA_TRY(cudaMemcpy(cnt.data(), ex->d_cellCnt + (size_t)slot * g.totalCells + L.cellBase, nCells * sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(list.data(), ex->d_cellList + (size_t)slot * g.cellListStride + L.cellListOff,
                        (size_t)nCells * L.cellCap * sizeof(unsigned), cudaMemcpyDeviceToHost));
  }
  int n = 0;
  for (int c = 0; c < nCells; ++c)
    for (int k = 0; k < cnt[c]; ++k, ++n)
      if (out && n < capacity) unpack_kp(list[(size_t)c * L.cellCap + k], &out[n]);
  *n_out = n;
  return ORBFE_OK;
}

int orbfe_debug_level_keypoints(orbfe_extractor* ex, int slot, int level, orbfe_keypoint* out, int capacity, int* n_out) {
  int rc = check_images(ex, slot, 1);
  if (rc) return rc;
  if (!ex->configured || level < 0 || level >= ex->g.nlevels || !n_out) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  CUDA_TRY(cudaSetDevice(ex->device));
  const Geom& g = ex->g;
  const LevelGeom& L = g.lv[level];
  int n = 0;
  std::vector<unsigned> v(L.outCap);
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  CUDA_TRY(cudaMemcpy(&n, ex->d_lvlCnt + (size_t)slot * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(v.data(), ex->d_lvlKp + (size_t)slot * g.totalOut + L.outOff, (size_t)L.outCap * sizeof(unsigned), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n && i < capacity && out; ++i) unpack_kp(v[i], &out[i]);
  *n_out = n;
  return ORBFE_OK;
}

int orbfe_debug_blurred(orbfe_extractor* ex, int slot, int level, uint8_t* dst, size_t dst_stride, int* w, int* h) {
  int rc = check_images(ex, slot, 1);
  if (rc) return rc;
  if (!ex->configured || level < 0 || level >= ex->g.nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  const LevelGeom& L = ex->g.lv[level];
  if (w) *w = L.w;
  if (h) *h = L.h;
  if (!dst) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(ex->device));
  CUDA_TRY(cudaMemcpy2DAsync(dst, dst_stride, ex->d_blur + (size_t)slot * ex->g.blurStride + L.blurOff, L.bpitch, L.w, L.h,
                             cudaMemcpyDeviceToHost, ex->stream));
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  return ORBFE_OK;
}

// ---- Frame::ComputeStereoMatches drop-in (frame.cpp:406-577) ----------------------------------
int orbfe_stereo_match(orbfe_extractor* left, orbfe_extractor* right, int n_left, const orbfe_keypoint* kps_left,
                       const uint8_t* desc_left, int n_right, const orbfe_keypoint* kps_right, const uint8_t* desc_right,
                       float bf, float baseline, float* u_right, float* depth, int* n_matched) {
  if (!left || !right || !u_right || !depth) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  if (n_left < 0 || n_right < 0 || (n_left && (!kps_left || !desc_left)) || (n_right && (!kps_right || !desc_right)))
    return orbfe_fail(ORBFE_ERR_INVALID, "bad keypoint arrays");
  if (!(baseline > 0.f) || !(bf > 0.f)) return orbfe_fail(ORBFE_ERR_INVALID, "bf and baseline must be positive");
  if (!left->configured || !right->configured) return orbfe_fail(ORBFE_ERR_INVALID, "extract both images first");
  if (left->device != right->device || left->g.w0 != right->g.w0 || left->g.h0 != right->g.h0 ||
      left->g.nlevels != right->g.nlevels || left->p.scale_factor != right->p.scale_factor)
    return orbfe_fail(ORBFE_ERR_INVALID, "left/right extractors differ in device, image size or pyramid");
  if (n_matched) *n_matched = 0;
  for (int i = 0; i < n_left; ++i) { u_right[i] = -1.0f; depth[i] = -1.0f; }
  if (n_left == 0 || n_right == 0) return ORBFE_OK;
  const Geom& g = left->g;
  if (n_left > g.totalOut || n_right > right->g.totalOut)
    return orbfe_fail(ORBFE_ERR_INVALID, "more keypoints than one image can produce (%d/%d > %d)", n_left, n_right, g.totalOut);
  // the kernels index the level geometry with keypoint.octave: a keypoint of another detector (octave -1, or more levels than
  // this handle has) must not reach them
  for (int i = 0; i < n_left; ++i)
    if (kps_left[i].octave < 0 || kps_left[i].octave >= g.nlevels)
      return orbfe_fail(ORBFE_ERR_INVALID, "left keypoint %d: octave %d outside [0, %d)", i, kps_left[i].octave, g.nlevels);
  for (int i = 0; i < n_right; ++i)
    if (kps_right[i].octave < 0 || kps_right[i].octave >= g.nlevels)
      return orbfe_fail(ORBFE_ERR_INVALID, "right keypoint %d: octave %d outside [0, %d)", i, kps_right[i].octave, g.nlevels);
  CUDA_TRY(cudaSetDevice(left->device));
  CUDA_TRY(cudaStreamSynchronize(right->stream));  // right pyramid complete
  CUDA_TRY(cudaStreamSynchronize(left->stream));
  orbfe_extractor* ex = left;
  const bool same = left == right;
  // the Frame's own arrays are authoritative (the reference reads keypoints_/descriptors_):
  // upload them into slot 0 of each handle's keypoint block (slot 1 of `left` if both are one handle)
  const size_t rslot = same ? 1 : 0;
  if (same && left->S < 2) return orbfe_fail(ORBFE_ERR_INVALID, "one shared handle needs max_images >= 2");
  // ... unless they ARE what the handles just produced (the usual case: keypoints_ / descriptors_ straight from the two
  // Compute calls): the pinned copies of the last download say so, and the device rows are already in place
  const size_t T = (size_t)g.totalOut;
  auto resident = [&](orbfe_extractor* e, size_t slot, int n, const orbfe_keypoint* k, const uint8_t* d) {
    return !e->slotDirty && e->h_n[slot] == n && !memcmp(e->h_kps + slot * T, k, (size_t)n * sizeof(orbfe_kp_dev)) &&
           !memcmp(e->h_desc + slot * T * 32, d, (size_t)n * 32);
  };
  int* cnts = ex->h_n + 2 * (size_t)ex->S + 1;
  cnts[0] = n_left; cnts[1] = n_right;
  if (!resident(left, 0, n_left, kps_left, desc_left)) {
    CUDA_TRY(cudaMemcpyAsync(left->d_kps, kps_left, (size_t)n_left * sizeof(orbfe_kp_dev), cudaMemcpyHostToDevice, ex->stream));
    CUDA_TRY(cudaMemcpyAsync(left->d_desc, desc_left, (size_t)n_left * 32, cudaMemcpyHostToDevice, ex->stream));
    CUDA_TRY(cudaMemcpyAsync(left->d_nKp, &cnts[0], sizeof(int), cudaMemcpyHostToDevice, ex->stream));
    left->slotDirty = true;
  }
  if (!resident(right, rslot, n_right, kps_right, desc_right)) {
    CUDA_TRY(cudaMemcpyAsync(right->d_kps + rslot * right->g.totalOut, kps_right, (size_t)n_right * sizeof(orbfe_kp_dev), cudaMemcpyHostToDevice, ex->stream));
    CUDA_TRY(cudaMemcpyAsync(right->d_desc + rslot * right->g.totalOut * 32, desc_right, (size_t)n_right * 32, cudaMemcpyHostToDevice, ex->stream));
    CUDA_TRY(cudaMemcpyAsync(right->d_nKp + rslot, &cnts[1], sizeof(int), cudaMemcpyHostToDevice, ex->stream));
    right->slotDirty = true;
  }
  StereoPair& P = ex->h_pairs[0];
  P.pyrL = left->d_pyr; P.pyrR = right->d_pyr + rslot * right->g.pyrStride;
  P.kpL = left->d_kps; P.kpR = right->d_kps + rslot * right->g.totalOut;
  P.descL = left->d_desc; P.descR = right->d_desc + rslot * right->g.totalOut * 32;
  P.nL = left->d_nKp; P.nR = right->d_nKp + rslot;
  P.uR = left->d_uR; P.depth = left->d_depth; P.sad = left->d_sad;
  P.rowStart = left->d_rowStart; P.rowItems = left->d_rowItems; P.rowCap = left->rowCap;
  ex->pairsCached = 0;
  CUDA_TRY(cudaMemcpyAsync(ex->d_pairs, ex->h_pairs, sizeof(StereoPair), cudaMemcpyHostToDevice, ex->stream));
  int rc = enqueue_stereo(ex, 1, bf, baseline);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(ex->h_uR, ex->d_uR, (size_t)n_left * sizeof(float), cudaMemcpyDeviceToHost, ex->stream));
  CUDA_TRY(cudaMemcpyAsync(ex->h_depth, ex->d_depth, (size_t)n_left * sizeof(float), cudaMemcpyDeviceToHost, ex->stream));
  CUDA_TRY(cudaMemcpyAsync(ex->h_n + ex->S, ex->d_nMatched, sizeof(int), cudaMemcpyDeviceToHost, ex->stream));
  CUDA_TRY(cudaStreamSynchronize(ex->stream));
  memcpy(u_right, ex->h_uR, (size_t)n_left * sizeof(float));
  memcpy(depth, ex->h_depth, (size_t)n_left * sizeof(float));
  if (n_matched) *n_matched = ex->h_n[ex->S];
  return ORBFE_OK;
}

}  // extern "C"
