/* Oracle: CPU restatement of MXNet's `_contrib_box_nms` and `_contrib_box_iou` in plain C.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED by the reference: the
 * operator lives in un-vendored, un-pinned Apache MXNet 1.4/1.5
 * (src/operator/contrib/bounding_box-inl.h, src/operator/tensor/sort_op.h); this file restates
 * its published algorithm.  The reference's call sites are
 *   models/definitions/yolo/yolo3.py:526-528 (x5) and yolo3_temporal.py:545-547   (box_nms)
 *   models/definitions/yolo/yolo_target.py:92                                     (box_iou)
 * Pinned by: the two upstream operator-doc examples (tests/test_oracle_kat.py); every assumption about the upstream operator is
 * listed with its known-answer test in oracle/ASSUMPTIONS.md; cross-checked on tie-free inputs against an independent
 * implementation (torchvision.ops.nms per class, tests/test_oracle_crosscheck.py).
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off; fp32 op order is part of the spec)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int32_t idx; } key_t_;

/* std::stable_sort(keys, '>') == order by (score desc, original index asc). */
static int cmp_desc_stable(const void* a, const void* b) {
    const key_t_* x = (const key_t_*)a; const key_t_* y = (const key_t_*)b;
    if (x->score > y->score) return -1;
    if (y->score > x->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* BoxNMS Intersect(): one axis, clamped at 0.  encode 0 = corner, 1 = center. */
static float intersect_1d(const float* a, const float* b, int encode) {
    float a1 = a[0], a2 = a[2], b1 = b[0], b2 = b[2];
    float left, right, w;
    if (encode == 1) {            /* center: (x, y, w, h) */
        float aw = a2 / 2.0f, bw = b2 / 2.0f;
        a2 = a1 + aw; a1 = a1 - aw; b2 = b1 + bw; b1 = b1 - bw;
    }
    left = a1 > b1 ? a1 : b1;
    right = a2 < b2 ? a2 : b2;
    w = right - left;
    return w > 0 ? w : 0;
}

/* BoxArea(): width * height, 0 when either extent is negative (upstream clamps degenerate boxes; oracle/ASSUMPTIONS.md A3).
 * encode 0: (x1,y1,x2,y2); 1: (x,y,w,h). */
static float box_area(const float* bx, int encode) {
    float w = encode == 0 ? bx[2] - bx[0] : bx[2];
    float h = encode == 0 ? bx[3] - bx[1] : bx[3];
    if (w < 0 || h < 0) return 0.0f;
    return w * h;
}

/* data (num_batch, num_elem, width) fp32 -> out same shape, record (num_batch, num_elem) int32
 * (original per-batch row index of each kept element, -1 elsewhere).
 * in_format/out_format: 0 corner, 1 center.  Returns 0. */
int ref_box_nms(const float* data, int64_t num_batch, int64_t num_elem, int width,
                float overlap_thresh, float valid_thresh, int topk, int coord_start,
                int score_index, int id_index, int background_id, int force_suppress,
                int in_format, int out_format, float* out, int32_t* record) {
    int64_t total = num_batch * num_elem;
    int64_t b, i;
    int64_t k_eff = (topk > 0 && topk < num_elem) ? topk : num_elem;
    for (i = 0; i < total * width; ++i) out[i] = -1.0f;
    if (record) for (i = 0; i < total; ++i) record[i] = -1;
    if (num_elem == 0) return 0;
    /* MXNet's Kernel<cpu>::Launch parallelises with OpenMP; images are independent. */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) private(i)
#endif
    for (b = 0; b < num_batch; ++b) {
        const float* in = data + b * num_elem * width;
        float* o = out + b * num_elem * width;
        int64_t nvalid = 0, n, r, p, row = 0;
        key_t_* keys = (key_t_*)malloc(sizeof(key_t_) * (size_t)num_elem);
        float* area = (float*)malloc(sizeof(float) * (size_t)(k_eff > 0 ? k_eff : 1));
        unsigned char* dead = (unsigned char*)malloc((size_t)(k_eff > 0 ? k_eff : 1));
        /* 1. valid filter, ascending index order (strict '>'; NaN invalid) */
        for (i = 0; i < num_elem; ++i) {
            float s = in[i * width + score_index];
            if (!(s > valid_thresh)) continue;
            if (id_index >= 0 && background_id >= 0 &&
                (int)in[i * width + id_index] == background_id) continue;
            keys[nvalid].score = s; keys[nvalid].idx = (int32_t)i; ++nvalid;
        }
        if (nvalid == 0) { free(keys); free(area); free(dead); continue; }
        /* 2. stable sort by score descending */
        qsort(keys, (size_t)nvalid, sizeof(key_t_), cmp_desc_stable);
        /* 3. only the first topk take part */
        n = nvalid < k_eff ? nvalid : k_eff;
        /* 4. areas */
        for (r = 0; r < n; ++r) {
            const float* bx = in + (int64_t)keys[r].idx * width + coord_start;
            area[r] = box_area(bx, in_format);
            dead[r] = 0;
        }
        /* 5. greedy suppression in rank order */
        for (r = 0; r + 1 < n; ++r) {
            const float* rb = in + (int64_t)keys[r].idx * width;
            if (dead[r]) continue;
            for (p = r + 1; p < n; ++p) {
                const float* pb = in + (int64_t)keys[p].idx * width;
                float inter, iou;
                if (dead[p]) continue;
                if (!force_suppress && id_index >= 0 &&
                    (int)rb[id_index] != (int)pb[id_index]) continue;
                inter = intersect_1d(rb + coord_start, pb + coord_start, in_format);
                inter *= intersect_1d(rb + coord_start + 1, pb + coord_start + 1, in_format);
                iou = inter / (area[r] + area[p] - inter);
                if (iou > overlap_thresh) dead[p] = 1;
            }
        }
        /* 6. compaction in rank order, all columns copied unmodified */
        for (r = 0; r < n; ++r) {
            if (dead[r]) continue;
            memcpy(o + row * width, in + (int64_t)keys[r].idx * width, sizeof(float) * (size_t)width);
            if (in_format != out_format) {
                float* c = o + row * width + coord_start;
                if (out_format == 0) {          /* center -> corner */
                    float x = c[0], y = c[1], w = c[2], h = c[3];
                    c[0] = x - w / 2.0f; c[1] = y - h / 2.0f; c[2] = x + w / 2.0f; c[3] = y + h / 2.0f;
                } else {                        /* corner -> center */
                    float l = c[0], t = c[1], rr = c[2], bb = c[3];
                    c[0] = (l + rr) / 2.0f; c[1] = (t + bb) / 2.0f; c[2] = rr - l; c[3] = bb - t;
                }
            }
            if (record) record[b * num_elem + row] = keys[r].idx;
            ++row;
        }
        free(keys); free(area); free(dead);
    }
    /* whole-input "no valid element" case == the -1 prefill above */
    return 0;
}

/* _contrib_box_iou(lhs (n,4), rhs (m,4), corner) -> (n,m): inter<=0 ? 0 : inter/(al+ar-inter). */
int ref_box_iou(const float* lhs, int64_t n, const float* rhs, int64_t m, float* out) {
    int64_t i, j;
    for (i = 0; i < n; ++i) {
        const float* a = lhs + 4 * i;
        float al = box_area(a, 0);
        for (j = 0; j < m; ++j) {
            const float* b = rhs + 4 * j;
            float ar = box_area(b, 0);
            float inter = intersect_1d(a, b, 0) * intersect_1d(a + 1, b + 1, 0);
            out[i * m + j] = (inter <= 0) ? 0.0f : inter / (al + ar - inter);
        }
    }
    return 0;
}
