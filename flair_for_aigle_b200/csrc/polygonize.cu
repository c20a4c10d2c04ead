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
#include <chrono>
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
static thread_local std::vector<RingStore> g_parts;   // the bands' rings between fz_trace_rings and fz_trace_rings_fetch

static void douglas_peucker(const std::vector<int32_t>& px, const std::vector<int32_t>& py, size_t a, size_t b, double tol2,
                            std::vector<uint8_t>& keep, std::vector<std::pair<size_t, size_t>>& stack) {
  // iterative DP on the open chain a..b (both kept); `stack` is the caller's scratch (no allocation per ring)
  stack.clear();
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
// Rings whose START lies in rows [y0, y1) of the label image.  A ring starts at its first east-heading top edge in row-major
// order, which is where a single sequential scan would pick it up; the bands are dealt to threads, so the unit of parallel work
// is the RING, not the component: one giant component (a background class with a million holes) spreads over all threads.
// A thread walks every ring it meets in its band once (its own one-bit-per-pixel "top edge walked" map), notes the smallest
// top-edge position on the way, and emits the ring only if that is the edge it started from -- otherwise the ring starts in
// an earlier band and belongs to that band's thread (which also walks it: rings crossing k bands are walked k times, in
// parallel; holes and small components, the bulk of the work, are walked once).  The owner's walk begins at the true start,
// so vertex order, and after the caller's merge by start position the ring order, are those of the sequential scan for any
// thread count.
static void trace_band(const int32_t* labels, int H, int W, const int32_t* keep_roots, int n_keep, double simplify_px,
                       int y0, int y1, RingStore& out) {
  RingStore rs;        // thread-local while tracing: the callers' stores sit side by side in one array, and growing vectors
                       // whose headers share cache lines from several threads costs more than the walks themselves
  rs.ring_offset.assign(1, 0);
  auto kept = [&](int32_t r) { return std::binary_search(keep_roots, keep_roots + n_keep, r); };
  auto lab = [&](int x, int y) -> int32_t {          // -1 outside the raster
    return (x < 0 || y < 0 || x >= W || y >= H) ? -1 : labels[static_cast<size_t>(y) * W + x];
  };
  // one bit per pixel: "the top edge of this pixel has been walked" (by THIS thread)
  std::vector<uint8_t> seen((static_cast<size_t>(H) * W + 7) / 8, 0);
  std::vector<int32_t> vx, vy, px, py;
  std::vector<uint8_t> keep;
  std::vector<std::pair<size_t, size_t>> dp_stack;
  const double tol2 = simplify_px * simplify_px;
  // direction d: 0 = east (+x), 1 = south (+y), 2 = west, 3 = north; component on the right-hand side
  static const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1};
  int32_t last_lab = -2;
  bool last_keep = false;
  const auto t_band = std::chrono::steady_clock::now();
  int64_t steps_mine = 0, steps_other = 0;
  for (int y = y0; y < y1; ++y)
    for (int x = 0; x < W; ++x) {
      const int32_t L = labels[static_cast<size_t>(y) * W + x];
      if (y > 0 && labels[static_cast<size_t>(y - 1) * W + x] == L) continue;      // no boundary above this pixel
      const size_t bit = static_cast<size_t>(y) * W + x;
      if (seen[bit >> 3] & (1u << (bit & 7))) continue;
      if (L != last_lab) {
        last_lab = L;
        last_keep = kept(L);
      }
      if (!last_keep) continue;
      // walk the ring that contains the east-heading top edge of (x, y), starting at vertex (x, y)
      vx.clear();
      vy.clear();
      int cx = x, cy = y, d = 0;
      int64_t area2 = 0;
      bool mine = true;                                   // until a top edge of this ring turns up BEFORE the one we started from
      do {
        if (d == 0) {
          const size_t b2 = static_cast<size_t>(cy) * W + cx;
          seen[b2 >> 3] |= static_cast<uint8_t>(1u << (b2 & 7));
          if (b2 < bit && mine) {                         // the ring starts in an earlier band: finish marking, emit nothing
            mine = false;
            vx.clear();
            vy.clear();
          }
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
          if (mine) {
            vx.push_back(cx);
            vy.push_back(cy);
          }
          d = nd;
        }
        ++(mine ? steps_mine : steps_other);
      } while (!(cx == x && cy == y && d == 0));
      if (!mine) continue;
      // vertices hold every corner once (the start vertex is a corner: nothing of L lies above-left on its ring)
      const size_t nv = vx.size();
      if (nv < 4) continue;
      keep.assign(nv + 1, 0);
      px.assign(vx.begin(), vx.end());
      py.assign(vy.begin(), vy.end());
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
        douglas_peucker(px, py, 0, far, tol2, keep, dp_stack);
        douglas_peucker(px, py, far, nv, tol2, keep, dp_stack);
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
  if (getenv("FZ_TRACE_TIMING"))
    fprintf(stderr, "[trace_band] rows [%d,%d): %.1f ms, %zu rings, steps on own rings %lld, on rings of earlier bands %lld\n", y0, y1,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_band).count(), rs.ring_root.size(),
            static_cast<long long>(steps_mine), static_cast<long long>(steps_other));
  out = std::move(rs);
}
}  // namespace fz

extern "C" int fz_trace_rings(const int32_t* labels, int H, int W, const int32_t* keep_roots, int n_keep,
                              double simplify_px, int64_t* n_rings, int64_t* n_points) {
  using namespace fz;
  FZ_REQUIRE(labels && H > 0 && W > 0 && n_keep >= 0 && n_rings && n_points, "fz_trace_rings: bad arguments");
  const auto t_begin = std::chrono::steady_clock::now();
  // host threads (FZ_TRACE_THREADS overrides), each with its own H*W-bit bitmap and its own band of rows; small rasters stay
  // on one thread
  int threads = static_cast<int>(std::thread::hardware_concurrency());
  if (const char* e = getenv("FZ_TRACE_THREADS")) threads = atoi(e);
  else if (static_cast<int64_t>(H) * W < (1 << 20)) threads = 1;
  threads = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
  if (threads > H) threads = H;
  std::vector<RingStore> parts(threads);
  if (threads == 1) {
    trace_band(labels, H, W, keep_roots, n_keep, simplify_px, 0, H, parts[0]);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
      const int y0 = static_cast<int>(static_cast<int64_t>(H) * t / threads), y1 = static_cast<int>(static_cast<int64_t>(H) * (t + 1) / threads);
      pool.emplace_back(trace_band, labels, H, W, keep_roots, n_keep, simplify_px, y0, y1, std::ref(parts[t]));
    }
    for (auto& th : pool) th.join();
  }
  const auto t_traced = std::chrono::steady_clock::now();
  // a band's rings all start inside the band, and the bands are ordered: concatenating the parts IS the order of one row-major
  // scan.  Nothing is merged here; fz_trace_rings_fetch copies every part to its place in the caller's arrays, in parallel.
  int64_t rings = 0, points = 0;
  for (const RingStore& p : parts) {
    rings += static_cast<int64_t>(p.ring_root.size());
    points += p.ring_offset.back();
  }
  g_parts = std::move(parts);
  *n_rings = rings;
  *n_points = points;
  if (getenv("FZ_TRACE_TIMING"))
    fprintf(stderr, "[fz_trace_rings] %d threads: walk %.1f ms\n", threads,
            std::chrono::duration<double, std::milli>(t_traced - t_begin).count());
  return 0;
}

extern "C" int fz_trace_rings_fetch(int32_t* ring_root, uint8_t* ring_is_hole, int64_t* ring_offset, double* xy) {
  using namespace fz;
  FZ_REQUIRE(ring_root && ring_is_hole && ring_offset && xy, "fz_trace_rings_fetch: null output");
  std::vector<RingStore> parts = std::move(g_parts);
  g_parts.clear();
  std::vector<int64_t> ring_base(parts.size() + 1, 0), point_base(parts.size() + 1, 0);
  for (size_t t = 0; t < parts.size(); ++t) {
    ring_base[t + 1] = ring_base[t] + static_cast<int64_t>(parts[t].ring_root.size());
    point_base[t + 1] = point_base[t] + parts[t].ring_offset.back();
  }
  ring_offset[0] = 0;
  auto copy_part = [&](size_t t) {
    const RingStore& p = parts[t];
    const int64_t rb = ring_base[t], pb = point_base[t];
    std::copy(p.ring_root.begin(), p.ring_root.end(), ring_root + rb);
    std::copy(p.ring_hole.begin(), p.ring_hole.end(), ring_is_hole + rb);
    for (size_t i = 0; i < p.ring_root.size(); ++i) ring_offset[rb + 1 + static_cast<int64_t>(i)] = pb + p.ring_offset[i + 1];
    std::copy(p.xy.begin(), p.xy.end(), xy + 2 * pb);
  };
  if (parts.size() <= 1) {
    for (size_t t = 0; t < parts.size(); ++t) copy_part(t);
  } else {
    std::vector<std::thread> pool;
    for (size_t t = 0; t < parts.size(); ++t) pool.emplace_back(copy_part, t);
    for (auto& th : pool) th.join();
  }
  return 0;
}
