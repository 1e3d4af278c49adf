/* CPU oracle for greedy IoU NMS - TEST INFRASTRUCTURE ONLY (see oracle/fce_oracle.py header).
 *
 * Restates the algorithm that ultralytics/utils/nms.py:151-154 dispatches to once torchvision is
 * imported: torchvision.ops.nms (torchvision 0.26.0, binary wheel - source not under
 * /root/reference; published algorithm: torchvision/csrc/ops/cpu/nms_kernel.cpp), which
 * ultralytics/utils/nms.py:239-296 (TorchNMS.nms) documents as semantically identical:
 *   - visit boxes by descending score, ties by ascending index (stable sort);
 *   - a visited, unsuppressed box i is kept and suppresses every later j with
 *     inter / (area_i + area_j - inter) > thr, all in fp32, no epsilon, no FMA contraction;
 *   - the threshold compare happens in double (the kernel's iou_threshold is a C double).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/build_oracle.py).
 */
#include <stdint.h>
#include <stdlib.h>

static const float *g_scores;
static int cmp_desc(const void *a, const void *b) {
    int32_t i = *(const int32_t *)a, j = *(const int32_t *)b;
    float si = g_scores[i], sj = g_scores[j];
    if (si > sj) return -1;
    if (si < sj) return 1;
    return (i > j) - (i < j);
}

/* boxes: n x 4 (x1,y1,x2,y2) fp32; returns number kept, indices (into the n inputs) in keep[] */
int fce_oracle_nms(const float *boxes, const float *scores, int n, double iou_thr, int64_t *keep) {
    if (n <= 0) return 0;
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * n);
    uint8_t *dead = (uint8_t *)calloc(n, 1);
    float *area = (float *)malloc(sizeof(float) * n);
    for (int i = 0; i < n; ++i) {
        order[i] = i;
        volatile float w = boxes[4 * i + 2] - boxes[4 * i + 0];
        volatile float h = boxes[4 * i + 3] - boxes[4 * i + 1];
        area[i] = w * h;
    }
    g_scores = scores;
    qsort(order, n, sizeof(int32_t), cmp_desc);
    int nk = 0;
    for (int a = 0; a < n; ++a) {
        int i = order[a];
        if (dead[i]) continue;
        keep[nk++] = i;
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float ia = area[i];
        for (int b = a + 1; b < n; ++b) {
            int j = order[b];
            if (dead[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            float w = xx2 - xx1; if (!(w > 0.f)) w = 0.f;
            float h = yy2 - yy1; if (!(h > 0.f)) h = 0.f;
            volatile float inter = w * h;
            volatile float uni = ia + area[j];
            uni = uni - inter;
            float ovr = inter / uni;
            if ((double)ovr > iou_thr) dead[j] = 1;
        }
    }
    free(order); free(dead); free(area);
    return nk;
}
