// The reference's own BCR interface (bcr.h:43-49) on top of the GPU builder (bcr.cu), so that `fermi ropebwt -a bcr`
// (ropebwt.c:47-158) links against libfermi_b200 UNMODIFIED in place of bcr.c: bcr_init / bcr_append / bcr_build / bcr_itr_init /
// bcr_itr_next / bcr_destroy and the bcr_verbose global, with the reference's argument meaning:
//   bcr_init(is_threaded, tmpfn)   both hints are irrelevant on the GPU (all cycles run there, sequences stay in host memory until
//                                  bcr_build) and ignored; the CUDA device is taken from FMG_DEVICE (default 0)
//   bcr_append(b, len, seq)        nt6 codes 1..4, 1 <= len < 65536 (asserted by the reference, bcr.c:361; an error here)
//   bcr_build(b)                   the BWT of the appended sequences; the process exits with a message if no GPU is present -- there
//                                  is no CPU path
//   bcr_itr_next(itr, &l)          the byte run-length stream (len << 3 | symbol, len <= 31, bcr.c:20-126) in chunks, NULL at the
//                                  end; the bytes differ from the reference's (its runs are split by bucket and 1 MB block
//                                  boundaries) but decode to the same BWT, which is all ropebwt.c:127-144 and rld.c:295-309 rely on
// The iterator is calloc'd because ropebwt.c:143 releases it with free().
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../../include/fermi_b200.h"

extern "C" {

int bcr_verbose = 2;                                      // bcr.c:10

struct bcr_s {
    fmg_bcr_t *h;
    uint8_t *rle;
    int64_t n_rle;
};
struct bcritr_s {
    const bcr_s *b;
    int64_t pos;
};
typedef struct bcr_s bcr_t;
typedef struct bcritr_s bcritr_t;

bcr_t *bcr_init(int is_threaded, const char *tmpfn) {
    (void)is_threaded; (void)tmpfn;
    const char *dev = std::getenv("FMG_DEVICE");
    bcr_t *b = static_cast<bcr_t *>(std::calloc(1, sizeof(bcr_t)));
    b->h = fmg_bcr_init(dev ? std::atoi(dev) : 0);
    return b;
}

void bcr_destroy(bcr_t *b) {
    if (!b) return;
    fmg_bcr_destroy(b->h);
    fmg_free(b->rle);
    std::free(b);
}

void bcr_append(bcr_t *b, int len, const uint8_t *seq) {
    if (len < 1 || len >= 65536 || fmg_bcr_append(b->h, len, seq) != 0) {
        std::fprintf(stderr, "[E::%s] sequences must hold 1..65535 bases A/C/G/T (bcr.c:361, ropebwt.c:98)\n", __func__);
        std::exit(1);
    }
}

void bcr_build(bcr_t *b) {
    const int old = fmg_verbose;
    if (bcr_verbose >= 3 && fmg_verbose < 3) fmg_verbose = 3;
    if (fmg_bcr_build(b->h) != 0 || fmg_bcr_rle(b->h, &b->rle, &b->n_rle) != 0) {
        std::fprintf(stderr, "[E::%s] the GPU build failed (libfermi_b200 has no CPU path)\n", __func__);
        std::exit(1);
    }
    fmg_verbose = old;
}

bcritr_t *bcr_itr_init(const bcr_t *b) {
    bcritr_t *itr = static_cast<bcritr_t *>(std::calloc(1, sizeof(bcritr_t)));
    itr->b = b; itr->pos = 0;
    return itr;
}

const uint8_t *bcr_itr_next(bcritr_t *itr, int *l) {
    const int64_t chunk = 1 << 20;                        // RLL_BLOCK_SIZE of the reference (bcr.c:24)
    if (itr->pos >= itr->b->n_rle) return nullptr;
    const uint8_t *s = itr->b->rle + itr->pos;
    const int64_t left = itr->b->n_rle - itr->pos;
    *l = (int)(left < chunk ? left : chunk);
    itr->pos += *l;
    return s;
}

} // extern "C"
