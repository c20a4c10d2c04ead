// Polygonisation of the class raster (inference.py:356-407: per class, rasterio.features.shapes(mask) with its default
// 4-connectivity -> one polygon per connected component -> area filter -> simplify).
//
// Device part: connected-component labelling of the uint8 class raster, 4-connectivity, all classes at once (two pixels
// are connected iff they are edge neighbours of the same class, which is what the per-class masks of the reference
// give).  Union-find with atomicMin on the parent array (label equivalence, Komura / Playne-Hawick style): every set's
// root is its smallest linear pixel index, so the labels do not depend on the order the unions happen in.
//   init    : parent = start of the pixel's horizontal same-class run inside its warp's 32-pixel span (ballot),
//             which removes 31/32 of the horizontal unions and keeps the chains short;
//   merge   : horizontal unions only where a run continues across a span border, vertical unions only where the
//             left neighbour does not already imply them;
//   compress: label = find(pixel).
// Then per-component pixel counts (warp-aggregated atomics) and a compacted component table.
// HBM traffic: 1 B/px class reads (x3 passes) + 4 B/px label writes/reads (x4) ~ 20 B/px algorithmic.
//
// Host part: ring tracing on the label image (boundary following on the pixel-corner lattice, component on the right,
// right turn first = 4-connectivity at saddle points), hole rings told from exterior rings by orientation, and a
// Douglas-Peucker pass per ring.  It runs on the CPU because it is pointer chasing over the boundary only
// (O(perimeter) after one O(H*W) scan), on host threads that each own a share of the components; the O(H*W) labelling is what the GPU takes over.
#include <algorithm>
#include <cmath>
#include <functional>
#include <thread>
#include <vector>

#include "common.h"
#include "../../include/flair_zonal_b200.h"

namespace fz {

__device__ __forceinline__ int ccl_find(const int* __restrict__ parent, int i) {
  int p = __ldcg(parent + i);
  while (p != i) {
    i = p;
    p = __ldcg(parent + i);
  }
  return i;
}

__device__ __forceinline__ void ccl_unite(int* parent, int a, int b) {
  bool done;
  do {
    a = ccl_find(parent, a);
    b = ccl_find(parent, b);
    if (a < b) {
      const int old = atomicMin(parent + b, a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const int old = atomicMin(parent + a, b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// grid (ceil(W/256), H), block 256: one thread per pixel, a warp = 32 consecutive pixels of a row
__global__ void __launch_bounds__(256) ccl_init_kernel(const uint8_t* __restrict__ cls, int* __restrict__ parent, int H,
                                                       int W) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const bool in = x < W;
  const int c = in ? cls[static_cast<size_t>(y) * W + x] : -1;
  const int left = __shfl_up_sync(0xffffffffu, c, 1);
  const bool start = !in || lane == 0 || left != c;
  const unsigned starts = __ballot_sync(0xffffffffu, start);
  if (!in) return;
  const unsigned below = starts & (0xffffffffu >> (31 - lane));   // run starts at lanes <= this one
  const int start_lane = 31 - __clz(below);
  parent[static_cast<size_t>(y) * W + x] = y * W + x - (lane - start_lane);
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint8_t* __restrict__ cls, int* __restrict__ parent, int H,
                                                        int W) {
  const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int i = y * W + x;
  const int c = cls[i];
  const bool same_left = x > 0 && cls[i - 1] == c;
  if (same_left && (threadIdx.x & 31) == 0) ccl_unite(parent, i, i - 1);      // run continues across the span border
  if (y > 0 && cls[i - W] == c) {
    // the left pixel's vertical union already joins the two rows when both left neighbours are of this class
    if (!(same_left && cls[i - W - 1] == c)) ccl_unite(parent, i, i - W);
  }
}

__global__ void __launch_bounds__(256) ccl_compress_kernel(int* __restrict__ parent, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  parent[i] = ccl_find(parent, i);
}

// area[root] += pixels; *n_roots += roots.  One atomic per distinct root per warp.
__global__ void __launch_bounds__(256) ccl_area_kernel(const int* __restrict__ labels, int* __restrict__ area,
                                                       int* __restrict__ n_roots, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const bool in = i < n;
  const unsigned active = __ballot_sync(0xffffffffu, in);
  if (!in) return;
  const int r = labels[i];
  const unsigned peers = __match_any_sync(active, r);
  const int lane = threadIdx.x & 31;
  if (lane == __ffs(peers) - 1) atomicAdd(area + r, __popc(peers));
  const unsigned roots = __ballot_sync(active, r == i);
  if (lane == __ffs(active) - 1 && roots) atomicAdd(n_roots, __popc(roots));
}

// compacted table of the components with area >= min_area_px and class != ignore_class, arbitrary order (the host sorts
// by root): rec[3k] = root, area, class; *counter ends as the number of such components (may exceed cap: call again)
__global__ void __launch_bounds__(256) ccl_table_kernel(const uint8_t* __restrict__ cls, const int* __restrict__ labels,
                                                        const int* __restrict__ area, int* __restrict__ counter,
                                                        int* __restrict__ rec, int cap, int n, int min_area_px,
                                                        int ignore_class) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n || labels[i] != i) return;
  if (area[i] < min_area_px || cls[i] == ignore_class) return;        // the reference's filters, applied before the download
  const int k = atomicAdd(counter, 1);
  if (k < cap) {
    rec[3 * k + 0] = i;
    rec[3 * k + 1] = area[i];
    rec[3 * k + 2] = cls[i];
  }
}

// ------------------------------------------------------------------------------------------------ host ring tracer
struct RingStore {
  std::vector<size_t> ring_start;     // linear index of the ring's first top edge (its position in a row-major scan)
  std::vector<int32_t> ring_root;
  std::vector<int64_t> ring_offset;   // n_rings + 1
  std::vector<uint8_t> ring_hole;
  std::vector<double> xy;
};
static thread_local RingStore g_rings;

static void douglas_peucker(const std::vector<int32_t>& px, const std::vector<int32_t>& py, size_t a, size_t b, double tol2,
                            std::vector<uint8_t>& keep) {
  // iterative DP on the open chain a..b (both kept)
  std::vector<std::pair<size_t, size_t>> stack;
  stack.emplace_back(a, b);
  while (!stack.empty()) {
    const auto [s, e] = stack.back();
    stack.pop_back();
    if (e <= s + 1) continue;
    const double ax = px[s], ay = py[s], bx = px[e], by = py[e];
    const double dx = bx - ax, dy = by - ay, len2 = dx * dx + dy * dy;
    double worst = -1.0;
    size_t wi = s;
    for (size_t i = s + 1; i < e; ++i) {
      double d2;
      if (len2 == 0.0) {
        d2 = (px[i] - ax) * (px[i] - ax) + (py[i] - ay) * (py[i] - ay);
      } else {
        double t = ((px[i] - ax) * dx + (py[i] - ay) * dy) / len2;
        t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
        const double qx = ax + t * dx - px[i], qy = ay + t * dy - py[i];
        d2 = qx * qx + qy * qy;
      }
      if (d2 > worst) {
        worst = d2;
        wi = i;
      }
    }
    if (worst > tol2) {
      keep[wi] = 1;
      stack.emplace_back(s, wi);
      stack.emplace_back(wi, e);
    }
  }
}

}  // namespace fz

extern "C" int fz_ccl_label(const uint8_t* raster, int32_t* labels, int H, int W, void* stream) {
  using namespace fz;
  FZ_REQUIRE(H > 0 && W > 0 && static_cast<int64_t>(H) * W < (1LL << 31), "fz_ccl_label: H=%d W=%d (H*W must fit int32)", H, W);
  FZ_REQUIRE(H <= 65535, "fz_ccl_label: H=%d exceeds the grid limit", H);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid((W + 255) / 256, H);
  const int n = H * W;
  ccl_init_kernel<<<grid, 256, 0, st>>>(raster, labels, H, W);
  ccl_merge_kernel<<<grid, 256, 0, st>>>(raster, labels, H, W);
  ccl_compress_kernel<<<(n + 255) / 256, 256, 0, st>>>(labels, n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_ccl_areas(const int32_t* labels, int32_t* area_zeroed, int32_t* n_roots_zeroed, int H, int W,
                            void* stream) {
  using namespace fz;
  FZ_REQUIRE(H > 0 && W > 0 && static_cast<int64_t>(H) * W < (1LL << 31), "fz_ccl_areas: bad shape");
  const int n = H * W;
  ccl_area_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(labels, area_zeroed, n_roots_zeroed, n);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fz_ccl_table(const uint8_t* raster, const int32_t* labels, const int32_t* area, int32_t* counter_zeroed,
                            int32_t* records, int capacity, int min_area_px, int ignore_class, int H, int W,
                            void* stream) {
  using namespace fz;
  FZ_REQUIRE(H > 0 && W > 0 && capacity >= 0, "fz_ccl_table: bad arguments");
  const int n = H * W;
  ccl_table_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(raster, labels, area, counter_zeroed,
                                                                                       records, capacity, n, min_area_px, ignore_class);
  FZ_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace fz {
// Rings of the components owned by thread `tid` of `n_threads` (components are dealt to threads by a hash of their label, so
// every ring -- exterior or hole -- is walked exactly once, by the thread that owns its component).  Each thread scans the
// whole label image (cheap next to the walks) and records where each ring starts; the caller merges the threads' rings by
// that position, which reproduces the order of a single row-major scan.
static void trace_band(const int32_t* labels, int H, int W, const int32_t* keep_roots, int n_keep, double simplify_px,
                       int tid, int n_threads, RingStore& rs) {
  rs.ring_offset.assign(1, 0);
  auto kept = [&](int32_t r) { return std::binary_search(keep_roots, keep_roots + n_keep, r); };
  auto lab = [&](int x, int y) -> int32_t {          // -1 outside the raster
    return (x < 0 || y < 0 || x >= W || y >= H) ? -1 : labels[static_cast<size_t>(y) * W + x];
  };
  // one bit per pixel: "the top edge of this pixel has been walked" (by THIS thread)
  std::vector<uint8_t> seen((static_cast<size_t>(H) * W + 7) / 8, 0);
  std::vector<int32_t> vx, vy;
  std::vector<uint8_t> keep;
  const double tol2 = simplify_px * simplify_px;
  // direction d: 0 = east (+x), 1 = south (+y), 2 = west, 3 = north; component on the right-hand side
  static const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1};
  int32_t last_lab = -2;
  bool last_keep = false;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int32_t L = labels[static_cast<size_t>(y) * W + x];
      if (y > 0 && labels[static_cast<size_t>(y - 1) * W + x] == L) continue;      // no boundary above this pixel
      const size_t bit = static_cast<size_t>(y) * W + x;
      if (seen[bit >> 3] & (1u << (bit & 7))) continue;
      if (L != last_lab) {
        last_lab = L;
        last_keep = (n_threads == 1 || static_cast<int>((static_cast<uint32_t>(L) * 2654435761u >> 16) % n_threads) == tid) &&
                    kept(L);
      }
      if (!last_keep) continue;
      // walk the ring that contains the east-heading top edge of (x, y), starting at vertex (x, y)
      vx.clear();
      vy.clear();
      int cx = x, cy = y, d = 0;
      int64_t area2 = 0;
      do {
        if (d == 0) {
          const size_t b2 = static_cast<size_t>(cy) * W + cx;
          seen[b2 >> 3] |= static_cast<uint8_t>(1u << (b2 & 7));
        }
        const int nx = cx + DX[d], ny = cy + DY[d];
        area2 += static_cast<int64_t>(cx) * ny - static_cast<int64_t>(nx) * cy;
        cx = nx;
        cy = ny;
        // the two pixels ahead of vertex (cx, cy) seen along d: ahead-right (AR) and ahead-left (AL)
        int arx, ary, alx, aly;
        switch (d) {
          case 0: arx = cx; ary = cy; alx = cx; aly = cy - 1; break;
          case 1: arx = cx - 1; ary = cy; alx = cx; aly = cy; break;
          case 2: arx = cx - 1; ary = cy - 1; alx = cx - 1; aly = cy; break;
          default: arx = cx; ary = cy - 1; alx = cx - 1; aly = cy - 1; break;
        }
        int nd;
        if (lab(arx, ary) != L) nd = (d + 1) & 3;              // hug the component: right turn first (4-connectivity)
        else if (lab(alx, aly) == L) nd = (d + 3) & 3;         // left turn
        else nd = d;
        if (nd != d) {
          vx.push_back(cx);
          vy.push_back(cy);
          d = nd;
        }
      } while (!(cx == x && cy == y && d == 0));
      // vertices hold every corner once (the start vertex is a corner: nothing of L lies above-left on its ring)
      const size_t nv = vx.size();
      if (nv < 4) continue;
      keep.assign(nv + 1, 0);
      std::vector<int32_t> px(vx), py(vy);
      px.push_back(vx[0]);
      py.push_back(vy[0]);
      keep[0] = keep[nv] = 1;
      if (simplify_px > 0.0) {
        // closed ring: split at the vertex farthest from the first one, then simplify the two chains
        size_t far = 0;
        double fd = -1.0;
        for (size_t i = 1; i < nv; ++i) {
          const double ddx = px[i] - px[0], ddy = py[i] - py[0], d2 = ddx * ddx + ddy * ddy;
          if (d2 > fd) {
            fd = d2;
            far = i;
          }
        }
        keep[far] = 1;
        douglas_peucker(px, py, 0, far, tol2, keep);
        douglas_peucker(px, py, far, nv, tol2, keep);
      } else {
        std::fill(keep.begin(), keep.end(), 1);
      }
      size_t emitted = 0;
      for (size_t i = 0; i <= nv; ++i)
        if (keep[i]) {
          rs.xy.push_back(px[i]);
          rs.xy.push_back(py[i]);
          ++emitted;
        }
      rs.ring_start.push_back(bit);
      rs.ring_root.push_back(L);
      rs.ring_hole.push_back(area2 < 0 ? 1 : 0);     // y grows downwards: exterior rings (component on the right) have area2 > 0
      rs.ring_offset.push_back(rs.ring_offset.back() + static_cast<int64_t>(emitted));
    }
}
}  // namespace fz

extern "C" int fz_trace_rings(const int32_t* labels, int H, int W, const int32_t* keep_roots, int n_keep,
                              double simplify_px, int64_t* n_rings, int64_t* n_points) {
  using namespace fz;
  FZ_REQUIRE(labels && H > 0 && W > 0 && n_keep >= 0 && n_rings && n_points, "fz_trace_rings: bad arguments");
  // host threads (FZ_TRACE_THREADS overrides), each with its own H*W-bit bitmap; small rasters stay on one thread
  int threads = static_cast<int>(std::thread::hardware_concurrency());
  if (const char* e = getenv("FZ_TRACE_THREADS")) threads = atoi(e);
  else if (static_cast<int64_t>(H) * W < (1 << 20)) threads = 1;
  threads = threads < 1 ? 1 : (threads > 32 ? 32 : threads);
  std::vector<RingStore> parts(threads);
  if (threads == 1) {
    trace_band(labels, H, W, keep_roots, n_keep, simplify_px, 0, 1, parts[0]);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
      pool.emplace_back(trace_band, labels, H, W, keep_roots, n_keep, simplify_px, t, threads, std::ref(parts[t]));
    for (auto& th : pool) th.join();
  }
  // merge by ring start = the order of one row-major scan (each part is already sorted)
  std::vector<std::pair<size_t, std::pair<int, int>>> order;            // (start, (part, ring index))
  for (int t = 0; t < threads; ++t)
    for (size_t i = 0; i < parts[t].ring_start.size(); ++i) order.push_back({parts[t].ring_start[i], {t, static_cast<int>(i)}});
  std::sort(order.begin(), order.end());
  RingStore& rs = g_rings;
  rs = RingStore();
  rs.ring_offset.assign(1, 0);
  for (const auto& o : order) {
    const RingStore& p = parts[o.second.first];
    const int i = o.second.second;
    const int64_t a0 = p.ring_offset[i], a1 = p.ring_offset[i + 1];
    rs.ring_root.push_back(p.ring_root[i]);
    rs.ring_hole.push_back(p.ring_hole[i]);
    rs.xy.insert(rs.xy.end(), p.xy.begin() + 2 * a0, p.xy.begin() + 2 * a1);
    rs.ring_offset.push_back(rs.ring_offset.back() + (a1 - a0));
  }
  *n_rings = static_cast<int64_t>(rs.ring_root.size());
  *n_points = rs.ring_offset.back();
  return 0;
}

extern "C" int fz_trace_rings_fetch(int32_t* ring_root, uint8_t* ring_is_hole, int64_t* ring_offset, double* xy) {
  using namespace fz;
  RingStore& rs = g_rings;
  FZ_REQUIRE(ring_root && ring_is_hole && ring_offset && xy, "fz_trace_rings_fetch: null output");
  std::copy(rs.ring_root.begin(), rs.ring_root.end(), ring_root);
  std::copy(rs.ring_hole.begin(), rs.ring_hole.end(), ring_is_hole);
  std::copy(rs.ring_offset.begin(), rs.ring_offset.end(), ring_offset);
  std::copy(rs.xy.begin(), rs.xy.end(), xy);
  rs = RingStore();
  return 0;
}
