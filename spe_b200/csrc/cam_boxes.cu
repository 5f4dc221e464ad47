// CAM -> pseudo ground-truth box on the device: engine.get_pseudo_label (engine.py:310-352) with cams_deit.resize_cam
// (cams_deit.py:9-14) and cams_deit.get_bboxes (:34-58), for every (image, present class) pair at once.
//
// Reference, per pair, on the host: D2H copy of the [h, w] map, cv2.resize (bilinear) to the image size, min-max normalise, * 255 ->
// uint8, cv2.threshold(TOZERO, int(cam_thr * max)), cv2.findContours(RETR_TREE), the contour of maximal cv2.contourArea, its
// cv2.boundingRect, -> cxcywh / image size.  Every step is restated bit-exactly:
//   * cv2.resize (this OpenCV build dispatches CV_32F INTER_LINEAR to IPP): source coordinate f = (d + 0.5) * (n_in / n_out) - 0.5 in
//     double, i0 = floor(f), weight = float(f - i0), clamped at both edges; horizontal pass then vertical pass, each
//     fma(s1 - s0, weight, s0) in fp32 (found by matching cv2 bit for bit, tests/golden/make_cam_fixture.py);
//   * normalisation: (v - min) / (max - min) * 255 in fp32, truncated to uint8; the mask is u8 > thr;
//   * contours: foreground is 8-connected, background 4-connected (Suzuki border following).  The contour of maximal area is always
//     the OUTER border of a top-level component, and cv2.contourArea of an outer border (polygon through the border-pixel centres)
//     equals  #(2x2 pixel cells with 4 pixels inside) + 1/2 #(cells with 3 inside)  over the component WITH ITS HOLES FILLED (checked
//     against cv2 on 200 random maps).  So: label foreground (8) and background (4) components with one union-find, build the nesting
//     tree from the left neighbour of every component's first pixel in raster order (exactly how the border follower assigns
//     parents), map every pixel to its top-level enclosing component, count cells, take the arg-max (ties: the component found last in
//     raster order, cv2's contour order is reverse raster), and emit its bounding box.
// The `_multi_boxes` variant (engine.py:356-398, cams_deit.get_multi_bboxes :61-97: every contour -- outer borders at any nesting level
// and hole borders -- above an area ratio) uses the same labelling: nested outer borders through ancestor chains, hole borders through a
// local shoelace sum over the crack edges of the filled hole.
#include "common.cuh"
#include <limits.h>

namespace {

constexpr int CB_THREADS = 256;

struct CamGeom {
    int B, C, h, w;            // cams f32 [B, C, h, w]
    int rows, cols;            // resized map: rows x cols  (the reference passes dsize = (H_img, W_img) to cv2.resize: rows = W_img, cols = H_img)
    int npix;
};

__device__ __forceinline__ void lin_coef(int d, int n_in, int n_out, int& i0, int& i1, float& fr) {
    const double f = ((double)d + 0.5) * ((double)n_in / (double)n_out) - 0.5;
    int i = (int)floor(f);
    float t = (float)(f - (double)i);
    if (i < 0) { i = 0; t = 0.f; }
    if (i >= n_in - 1) { i = n_in - 1; t = 0.f; }
    i0 = i;
    i1 = min(i + 1, n_in - 1);
    fr = t;
}
__device__ __forceinline__ float lerp_ipp(float s0, float s1, float t) { return __fmaf_rn(__fsub_rn(s1, s0), t, s0); }

__device__ __forceinline__ float resized_value(const float* __restrict__ cam, const CamGeom& g, int y, int x) {
    int y0, y1, x0, x1;
    float fy, fx;
    lin_coef(y, g.h, g.rows, y0, y1, fy);
    lin_coef(x, g.w, g.cols, x0, x1, fx);
    const float r0 = lerp_ipp(cam[y0 * g.w + x0], cam[y0 * g.w + x1], fx);
    const float r1 = lerp_ipp(cam[y1 * g.w + x0], cam[y1 * g.w + x1], fx);
    return lerp_ipp(r0, r1, fy);
}

// order-preserving float <-> int (atomicMin / atomicMax on floats of either sign)
__device__ __forceinline__ int f2ord(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

// per map: minmax[2] (ordered ints), then the per-pixel arrays
struct CamWs {
    int* minmax;               // [npairs, 2]
    unsigned char* mask;       // [npairs, npix]
    int* label;                // [npairs, npix]  union-find parent; after flatten: root (= first pixel of the component in raster order)
    int* aux;                  // [npairs, npix]  bg roots: 1 = touches the image border; later: top-level component of every root (-1 = outside)
    int* parent;               // [npairs, npix]  nesting-tree parent of a root (-1 = outside / top level)
    int* area2;                // [npairs, npix]  2 x contour area of top-level fg roots
    int* bbox;                 // [npairs, 4, npix] min x, min y, max x, max y of fg roots
};

__global__ void __launch_bounds__(CB_THREADS) cam_minmax_kernel(const float* __restrict__ cams, const int* __restrict__ pairs, CamGeom g, CamWs ws) {
    __shared__ float smin[32], smax[32];
    const int m = blockIdx.y;
    const float* cam = cams + ((long long)pairs[2 * m] * g.C + pairs[2 * m + 1]) * g.h * g.w;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = blockIdx.x * CB_THREADS + threadIdx.x; i < g.npix; i += gridDim.x * CB_THREADS) {
        const float v = resized_value(cam, g, i / g.cols, i % g.cols);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    mn = -warp_max(-mn);
    mx = warp_max(mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { smin[warp] = mn; smax[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < CB_THREADS / 32; ++k) { mn = fminf(mn, smin[k]); mx = fmaxf(mx, smax[k]); }
        atomicMin(&ws.minmax[2 * m], f2ord(mn));
        atomicMax(&ws.minmax[2 * m + 1], f2ord(mx));
    }
}

__global__ void __launch_bounds__(CB_THREADS) cam_mask_init_kernel(const float* __restrict__ cams, const int* __restrict__ pairs, CamGeom g, CamWs ws, int thr_u8) {
    const int m = blockIdx.y;
    const float* cam = cams + ((long long)pairs[2 * m] * g.C + pairs[2 * m + 1]) * g.h * g.w;
    const float mn = ord2f(ws.minmax[2 * m]);
    const float span = __fsub_rn(ord2f(ws.minmax[2 * m + 1]), mn);                 // (cam - cam.min()).max()
    const long long base = (long long)m * g.npix;
    for (int i = blockIdx.x * CB_THREADS + threadIdx.x; i < g.npix; i += gridDim.x * CB_THREADS) {
        const float v = resized_value(cam, g, i / g.cols, i % g.cols);
        const float q = __fmul_rn(__fdiv_rn(__fsub_rn(v, mn), span), 255.f);
        const int u = (int)q;                                                     // astype(uint8) of a value in [0, 255]
        ws.mask[base + i] = (u & 255) > thr_u8 ? 1 : 0;
        ws.label[base + i] = i;
        ws.aux[base + i] = 0;
        ws.parent[base + i] = -1;
        ws.area2[base + i] = 0;
        ws.bbox[(base * 4) + 0LL * g.npix + i] = INT_MAX;
        ws.bbox[(base * 4) + 1LL * g.npix + i] = INT_MAX;
        ws.bbox[(base * 4) + 2LL * g.npix + i] = -1;
        ws.bbox[(base * 4) + 3LL * g.npix + i] = -1;
    }
}

__device__ __forceinline__ int uf_find(const int* L, int i) {
    while (true) {
        const int p = __ldcg(L + i);
        if (p == i) return i;
        i = p;
    }
}
__device__ __forceinline__ void uf_unite(int* L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(L + b, a);          // roots point to themselves: linking the larger root under the smaller one
        if (old == b) return;
        b = old;
    }
}

__global__ void __launch_bounds__(CB_THREADS) cam_merge_kernel(CamGeom g, CamWs ws) {
    const int m = blockIdx.y;
    const unsigned char* M = ws.mask + (long long)m * g.npix;
    int* L = ws.label + (long long)m * g.npix;
    for (int i = blockIdx.x * CB_THREADS + threadIdx.x; i < g.npix; i += gridDim.x * CB_THREADS) {
        const int y = i / g.cols, x = i % g.cols;
        const int v = M[i];
        if (x > 0 && M[i - 1] == v) uf_unite(L, i, i - 1);
        if (y > 0 && M[i - g.cols] == v) uf_unite(L, i, i - g.cols);
        if (v) {                                                                  // foreground: 8-connected
            if (y > 0 && x > 0 && M[i - g.cols - 1]) uf_unite(L, i, i - g.cols - 1);
            if (y > 0 && x + 1 < g.cols && M[i - g.cols + 1]) uf_unite(L, i, i - g.cols + 1);
        }
    }
}

__global__ void __launch_bounds__(CB_THREADS) cam_flatten_kernel(CamGeom g, CamWs ws) {
    const int m = blockIdx.y;
    const long long base = (long long)m * g.npix;
    const unsigned char* M = ws.mask + base;
    int* L = ws.label + base;
    for (int i = blockIdx.x * CB_THREADS + threadIdx.x; i < g.npix; i += gridDim.x * CB_THREADS) {
        const int r = uf_find(L, i);
        L[i] = r;                                   // only shortens paths: concurrent finds stay valid
        const int y = i / g.cols, x = i % g.cols;
        if (M[i]) {
            atomicMin(&ws.bbox[base * 4 + 0LL * g.npix + r], x);
            atomicMin(&ws.bbox[base * 4 + 1LL * g.npix + r], y);
            atomicMax(&ws.bbox[base * 4 + 2LL * g.npix + r], x);
            atomicMax(&ws.bbox[base * 4 + 3LL * g.npix + r], y);
        } else if (x == 0 || y == 0 || x == g.cols - 1 || y == g.rows - 1) {
            ws.aux[base + r] = 1;                   // this background region is the outside (cv2 pads the image with background)
        }
    }
}

// nesting tree: parent of a root = the region left of its first pixel (the border follower's rule)
__global__ void __launch_bounds__(CB_THREADS) cam_parent_kernel(CamGeom g, CamWs ws) {
    const int m = blockIdx.y;
    const long long base = (long long)m * g.npix;
    const unsigned char* M = ws.mask + base;
    const int* L = ws.label + base;
    for (int i = blockIdx.x * CB_THREADS + threadIdx.x; i < g.npix; i += gridDim.x * CB_THREADS) {
        if (L[i] != i) continue;
        const int x = i % g.cols;
        int par = -1;
        if (M[i]) {
            if (x > 0) {                            // left neighbour is background (else i would not be the first pixel)
                const int rb = L[i - 1];
                par = ws.aux[base + rb] ? -1 : rb;
            }
        } else if (!ws.aux[base + i]) {
            par = L[i - 1];                         // a hole: its left neighbour belongs to the enclosing foreground component
        }
        ws.parent[base + i] = par;
    }
}

// contour area of EVERY foreground component (any nesting level): a pixel is inside component X's outer border iff X is on the
// ancestor chain of the pixel's region in the nesting tree.  A 2x2 cell adds 1 (all four corners inside) or 1/2 (three) to X.
constexpr int CB_MAXCHAIN = 12;     // fg nodes tracked per cell (nesting deeper than 12 levels is ignored)
__global__ void __launch_bounds__(CB_THREADS) cam_cells_kernel(CamGeom g, CamWs ws) {
    const int m = blockIdx.y;
    const long long base = (long long)m * g.npix;
    const unsigned char* M = ws.mask + base;
    const int* L = ws.label + base;
    const int* P = ws.parent + base;
    const int ncell = (g.rows - 1) * (g.cols - 1);
    for (int c = blockIdx.x * CB_THREADS + threadIdx.x; c < ncell; c += gridDim.x * CB_THREADS) {
        const int y = c / (g.cols - 1), x = c % (g.cols - 1);
        const int i = y * g.cols + x;
        const int r[4] = {L[i], L[i + 1], L[i + g.cols], L[i + g.cols + 1]};
        if (r[0] == r[1] && r[0] == r[2] && r[0] == r[3]) {                // uniform cell (the common case): whole chain gets a full cell
            for (int q = r[0]; q >= 0; q = P[q])
                if (M[q]) atomicAdd(&ws.area2[base + q], 2);
            continue;
        }
        int node[CB_MAXCHAIN], cnt[CB_MAXCHAIN], n = 0;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            for (int q = r[k]; q >= 0; q = P[q]) {
                if (!M[q]) continue;
                int t = 0;
                while (t < n && node[t] != q) ++t;
                if (t == n) {
                    if (n == CB_MAXCHAIN) continue;
                    node[n] = q; cnt[n] = 0; ++n;
                }
                ++cnt[t];
            }
        }
        for (int t = 0; t < n; ++t) {
            if (cnt[t] == 4) atomicAdd(&ws.area2[base + node[t]], 2);
            else if (cnt[t] == 3) atomicAdd(&ws.area2[base + node[t]], 1);
        }
    }
}

// hole borders (cv2 traces them on the foreground pixels around the hole): 2 x signed polygon area by the shoelace formula summed
// over the crack edges of the FILLED hole region R (hole + everything inside it), walked clockwise.  Vertex of an edge = the outside
// pixel across it; the next edge follows from the two pixels ahead (convex corner: same R pixel, next side; concave: the diagonal
// pixel's side, same outside pixel; straight).  Every term is local, so no traversal order is needed.  Checked against cv2 on 19 000
// hole contours incl. salt-and-pepper masks (tests/golden/make_cam_fixture.py).  bbox of a hole contour = bbox of the vertices.
__device__ __forceinline__ bool in_chain(const int* P, int root, int X) {
    for (int q = root; q >= 0; q = P[q])
        if (q == X) return true;
    return false;
}
__global__ void __launch_bounds__(CB_THREADS) cam_hole_edges_kernel(CamGeom g, CamWs ws) {
    const int m = blockIdx.y;
    const long long base = (long long)m * g.npix;
    const unsigned char* M = ws.mask + base;
    const int* L = ws.label + base;
    const int* P = ws.parent + base;
    const int DX[4] = {0, 1, 0, -1}, DY[4] = {-1, 0, 1, 0};                // N, E, S, W
    for (int i = blockIdx.x * CB_THREADS + threadIdx.x; i < g.npix; i += gridDim.x * CB_THREADS) {
        if (M[i]) continue;
        const int Hh = L[i];
        if (P[Hh] < 0) continue;                                            // background connected to the image border: not a hole
        const int y = i / g.cols, x = i % g.cols;                           // holes never touch the border: all 8 neighbours exist
        unsigned int sum = 0;                                               // modular arithmetic: the partial sums exceed 32 bits, the total does not
        int x0 = INT_MAX, y0 = INT_MAX, x1 = -1, y1 = -1;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int ox = x + DX[d], oy = y + DY[d];
            if (in_chain(P, L[oy * g.cols + ox], Hh)) continue;             // neighbour inside R (hole pixel or an island in it)
            const int tx = DX[(d + 1) & 3], ty = DY[(d + 1) & 3];           // travel direction along this edge
            const int bx = x + tx, by = y + ty, ax = bx + DX[d], ay = by + DY[d];
            int nx, ny;
            if (!in_chain(P, L[by * g.cols + bx], Hh)) { nx = bx; ny = by; }            // convex corner
            else if (in_chain(P, L[ay * g.cols + ax], Hh)) { nx = ox; ny = oy; }        // concave corner
            else { nx = ax; ny = ay; }                                                  // straight
            sum += (unsigned int)(ox * ny - nx * oy);
            x0 = min(x0, ox); y0 = min(y0, oy); x1 = max(x1, ox); y1 = max(y1, oy);
        }
        if (x1 >= 0) {
            if (sum) atomicAdd(reinterpret_cast<unsigned int*>(&ws.area2[base + Hh]), sum);
            atomicMin(&ws.bbox[base * 4 + 0LL * g.npix + Hh], x0);
            atomicMin(&ws.bbox[base * 4 + 1LL * g.npix + Hh], y0);
            atomicMax(&ws.bbox[base * 4 + 2LL * g.npix + Hh], x1);
            atomicMax(&ws.bbox[base * 4 + 3LL * g.npix + Hh], y1);
        }
    }
}

// one block per map: arg-max of area2 over top-level foreground roots (ties: largest root index), -> cxcywh / image size
__global__ void __launch_bounds__(CB_THREADS) cam_select_kernel(CamGeom g, CamWs ws, float norm_x, float norm_y, float* __restrict__ boxes, int* __restrict__ xyxy) {
    __shared__ long long best[CB_THREADS];
    const int m = blockIdx.x;
    const long long base = (long long)m * g.npix;
    const unsigned char* M = ws.mask + base;
    const int* L = ws.label + base;
    long long b = -1;
    for (int i = threadIdx.x; i < g.npix; i += CB_THREADS) {
        if (L[i] == i && M[i] && ws.parent[base + i] < 0) {
            const long long key = ((long long)ws.area2[base + i] << 32) | (unsigned int)i;
            b = key > b ? key : b;
        }
    }
    best[threadIdx.x] = b;
    __syncthreads();
    for (int s = CB_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s && best[threadIdx.x + s] > best[threadIdx.x]) best[threadIdx.x] = best[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int x0 = 0, y0 = 0, x1 = 1, y1 = 1;                           // no contour: [0, 0, 1, 1] (cams_deit.py:56)
        if (best[0] >= 0) {
            const int r = (int)(best[0] & 0xffffffffLL);
            x0 = ws.bbox[base * 4 + 0LL * g.npix + r];
            y0 = ws.bbox[base * 4 + 1LL * g.npix + r];
            x1 = ws.bbox[base * 4 + 2LL * g.npix + r] + 1;            // boundingRect: [x, y, x + w, y + h]
            y1 = ws.bbox[base * 4 + 3LL * g.npix + r] + 1;
        }
        if (xyxy) { xyxy[4 * m] = x0; xyxy[4 * m + 1] = y0; xyxy[4 * m + 2] = x1; xyxy[4 * m + 3] = y1; }
        // box_xyxy_to_cxcywh on the integer box, then / [w, h, w, h]  (engine.py:311-320)
        boxes[4 * m + 0] = __fdiv_rn((float)(x0 + x1) * 0.5f, norm_x);
        boxes[4 * m + 1] = __fdiv_rn((float)(y0 + y1) * 0.5f, norm_y);
        boxes[4 * m + 2] = __fdiv_rn((float)(x1 - x0), norm_x);
        boxes[4 * m + 3] = __fdiv_rn((float)(y1 - y0), norm_y);
    }
}

// get_multi_bboxes (cams_deit.py:61-97): every contour (outer borders at any nesting level AND hole borders) whose area is
// >= area_ratio * the largest area, by decreasing area (equal areas: later raster position first -- cv2's tree order may differ for
// exact ties), each as boundingRect -> [x, y, x + w, y + h].  One block per map.
constexpr int CB_MAXBOX = 64;
__global__ void __launch_bounds__(CB_THREADS) cam_select_multi_kernel(CamGeom g, CamWs ws, float norm_x, float norm_y, double ratio, int max_boxes,
                                                                       float* __restrict__ boxes, int* __restrict__ xyxy, int* __restrict__ counts) {
    __shared__ int red[CB_THREADS];
    __shared__ int s_n, s_area[CB_MAXBOX], s_idx[CB_MAXBOX];
    const int m = blockIdx.x;
    const long long base = (long long)m * g.npix;
    const unsigned char* M = ws.mask + base;
    const int* L = ws.label + base;
    const int* P = ws.parent + base;
    int best = -1;
    for (int i = threadIdx.x; i < g.npix; i += CB_THREADS) {
        if (L[i] != i || (!M[i] && P[i] < 0)) continue;                    // not a root, or the outside background
        const int a = ws.area2[base + i];
        best = max(best, a < 0 ? -a : a);
    }
    red[threadIdx.x] = best;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int s = CB_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] = max(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    const int amax = red[0];
    const double thr = (double)amax * 0.5 * ratio;                          // areas[idx] >= areas[area_idx[0]] * area_ratio, in double as in Python
    for (int i = threadIdx.x; i < g.npix; i += CB_THREADS) {
        if (L[i] != i || (!M[i] && P[i] < 0)) continue;
        int a = ws.area2[base + i];
        a = a < 0 ? -a : a;
        if ((double)a * 0.5 >= thr) {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < CB_MAXBOX) { s_area[slot] = a; s_idx[slot] = i; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = min(s_n, CB_MAXBOX);
        for (int a = 1; a < n; ++a) {                                       // insertion sort: area descending, then root index descending
            const int ka = s_area[a], ki = s_idx[a];
            int b = a - 1;
            while (b >= 0 && (s_area[b] < ka || (s_area[b] == ka && s_idx[b] < ki))) { s_area[b + 1] = s_area[b]; s_idx[b + 1] = s_idx[b]; --b; }
            s_area[b + 1] = ka; s_idx[b + 1] = ki;
        }
        n = min(n, max_boxes);
        if (amax < 0) n = 0;
        const long long o = (long long)m * max_boxes;
        for (int k = 0; k < max(n, 1); ++k) {
            int x0 = 0, y0 = 0, x1 = 1, y1 = 1;                             // no contour: [[0, 0, 1, 1]] (cams_deit.py:95)
            if (n > 0) {
                const int r = s_idx[k];
                x0 = ws.bbox[base * 4 + 0LL * g.npix + r];
                y0 = ws.bbox[base * 4 + 1LL * g.npix + r];
                x1 = ws.bbox[base * 4 + 2LL * g.npix + r] + 1;
                y1 = ws.bbox[base * 4 + 3LL * g.npix + r] + 1;
            }
            if (xyxy) { xyxy[4 * (o + k)] = x0; xyxy[4 * (o + k) + 1] = y0; xyxy[4 * (o + k) + 2] = x1; xyxy[4 * (o + k) + 3] = y1; }
            boxes[4 * (o + k) + 0] = __fdiv_rn((float)(x0 + x1) * 0.5f, norm_x);
            boxes[4 * (o + k) + 1] = __fdiv_rn((float)(y0 + y1) * 0.5f, norm_y);
            boxes[4 * (o + k) + 2] = __fdiv_rn((float)(x1 - x0), norm_x);
            boxes[4 * (o + k) + 3] = __fdiv_rn((float)(y1 - y0), norm_y);
        }
        counts[m] = max(n, 1);
    }
}

__global__ void cam_minmax_init_kernel(int* minmax, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        minmax[2 * i] = INT_MAX;
        minmax[2 * i + 1] = INT_MIN;
    }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" __attribute__((visibility("default"))) int64_t spe_cam_boxes_workspace_bytes(int npairs, int rows, int cols) {
    const size_t npix = (size_t)rows * cols, n = (size_t)npairs;
    return (int64_t)(align256(n * 2 * 4) + align256(n * npix) + 4 * align256(n * npix * 4) + align256(n * npix * 16));
}

namespace {
int cam_boxes_run(const float* cams, int B, int C, int h, int w, const int32_t* pairs, int npairs, int rows, int cols, int thr_u8, float norm_x, float norm_y,
                  double area_ratio, int max_boxes, float* boxes_out, int32_t* xyxy_out, int32_t* counts_out, void* workspace, int64_t workspace_bytes, void* stream) {
    SPE_CHECK(cams && pairs && boxes_out && workspace && B > 0 && C > 0 && h > 0 && w > 0 && npairs > 0 && rows > 1 && cols > 1, "spe_cam_boxes: bad argument");
    SPE_CHECK((long long)rows * cols < (1LL << 30) && thr_u8 >= 0 && thr_u8 <= 255 && norm_x > 0.f && norm_y > 0.f, "spe_cam_boxes: bad argument");
    SPE_CHECK(workspace_bytes >= spe_cam_boxes_workspace_bytes(npairs, rows, cols), "spe_cam_boxes: workspace too small");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CamGeom g{B, C, h, w, rows, cols, rows * cols};
    const size_t npix = (size_t)g.npix, n = (size_t)npairs;
    char* p = reinterpret_cast<char*>(workspace);
    CamWs ws;
    ws.minmax = reinterpret_cast<int*>(p); p += align256(n * 2 * 4);
    ws.mask = reinterpret_cast<unsigned char*>(p); p += align256(n * npix);
    ws.label = reinterpret_cast<int*>(p); p += align256(n * npix * 4);
    ws.aux = reinterpret_cast<int*>(p); p += align256(n * npix * 4);
    ws.parent = reinterpret_cast<int*>(p); p += align256(n * npix * 4);
    ws.area2 = reinterpret_cast<int*>(p); p += align256(n * npix * 4);
    ws.bbox = reinterpret_cast<int*>(p);
    int gx = (g.npix + CB_THREADS - 1) / CB_THREADS;
    const int cap = (spe_num_sms() * 8 + npairs - 1) / npairs;
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    const dim3 grid(gx, npairs);
    cam_minmax_init_kernel<<<(npairs + 127) / 128, 128, 0, st>>>(ws.minmax, npairs);
    SPE_LAUNCHED();
    cam_minmax_kernel<<<grid, CB_THREADS, 0, st>>>(cams, pairs, g, ws);
    SPE_LAUNCHED();
    cam_mask_init_kernel<<<grid, CB_THREADS, 0, st>>>(cams, pairs, g, ws, thr_u8);
    SPE_LAUNCHED();
    cam_merge_kernel<<<grid, CB_THREADS, 0, st>>>(g, ws);
    SPE_LAUNCHED();
    cam_flatten_kernel<<<grid, CB_THREADS, 0, st>>>(g, ws);
    SPE_LAUNCHED();
    cam_parent_kernel<<<grid, CB_THREADS, 0, st>>>(g, ws);
    SPE_LAUNCHED();
    cam_cells_kernel<<<grid, CB_THREADS, 0, st>>>(g, ws);
    SPE_LAUNCHED();
    if (max_boxes <= 0) {
        cam_select_kernel<<<npairs, CB_THREADS, 0, st>>>(g, ws, norm_x, norm_y, boxes_out, xyxy_out);
        SPE_LAUNCHED();
    } else {
        cam_hole_edges_kernel<<<grid, CB_THREADS, 0, st>>>(g, ws);
        SPE_LAUNCHED();
        cam_select_multi_kernel<<<npairs, CB_THREADS, 0, st>>>(g, ws, norm_x, norm_y, area_ratio, max_boxes, boxes_out, xyxy_out, counts_out);
        SPE_LAUNCHED();
    }
    return 0;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int spe_cam_boxes(const float* cams, int B, int C, int h, int w, const int32_t* pairs, int npairs, int rows, int cols,
                                                                    int thr_u8, float norm_x, float norm_y, float* boxes_out, int32_t* xyxy_out, void* workspace,
                                                                    int64_t workspace_bytes, void* stream) {
    return cam_boxes_run(cams, B, C, h, w, pairs, npairs, rows, cols, thr_u8, norm_x, norm_y, 0.0, 0, boxes_out, xyxy_out, nullptr, workspace, workspace_bytes, stream);
}

extern "C" __attribute__((visibility("default"))) int spe_cam_boxes_multi(const float* cams, int B, int C, int h, int w, const int32_t* pairs, int npairs, int rows,
                                                                          int cols, int thr_u8, float norm_x, float norm_y, double area_ratio, int max_boxes,
                                                                          float* boxes_out, int32_t* xyxy_out, int32_t* counts_out, void* workspace,
                                                                          int64_t workspace_bytes, void* stream) {
    SPE_CHECK(max_boxes >= 1 && max_boxes <= CB_MAXBOX && counts_out && area_ratio >= 0.0, "spe_cam_boxes_multi: 1 <= max_boxes <= %d", CB_MAXBOX);
    return cam_boxes_run(cams, B, C, h, w, pairs, npairs, rows, cols, thr_u8, norm_x, norm_y, area_ratio, max_boxes, boxes_out, xyxy_out, counts_out, workspace,
                         workspace_bytes, stream);
}
