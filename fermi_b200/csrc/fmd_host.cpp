// Host-side .fmd reader / writer / encoder of libfermi_b200 (product code).
// File format and bit layout follow the reference so that files are interchangeable:
//   rld_dump rld.c:242-263, rld_restore rld.c:265-325, block/run layout rld.c:47-53,111-173.
#include "fmd_host.hpp"
#include "../../include/fermi_b200.h"
#include <cstring>
#include <cstdlib>
#include <memory>

int fmg_verbose = 3;
// FMG_VERBOSE in the environment overrides the default when the library is loaded (fm_verbose has no such hook: utils.c:8)
static const int fmg_verbose_from_env = [] { const char *e = std::getenv("FMG_VERBOSE"); if (e && *e) fmg_verbose = std::atoi(e); return 0; }();

namespace fmg {

static inline int ilog2_32(uint32_t v) { return v ? 31 - __builtin_clz(v) : -1; }

void FmdImage::finish_counts() {
    cnt[0] = 0;
    for (int c = 1; c <= kAlphabet; ++c) cnt[c] = cnt[c - 1] + mcnt[c];
    cnt[7] = cnt[6];
    mcnt[0] = cnt[6];
}

void FmdImage::build_frames() {
    const uint64_t n_blks = n_bytes * 8 / 64 / kBlockWords + 1;
    const uint64_t last = n_blocks() * kBlockWords;
    ibits = ilog2_32((uint32_t)(mcnt[0] / n_blks)) + 4;
    n_frames = ((mcnt[0] + (1ull << ibits) - 1) >> ibits) + 1;
    frame.assign(n_frames * 7, 0);
    uint64_t acc[6] = {0, 0, 0, 0, 0, 0}, k = 1;
    for (uint64_t i = kBlockWords; i <= last; i += kBlockWords) {
        if (header_is32(words[i])) {
            const uint32_t *h = reinterpret_cast<const uint32_t *>(&words[i]);
            for (int c = 0; c < 6; ++c) acc[c] += h[c + 1];
        } else {
            const uint16_t *h = reinterpret_cast<const uint16_t *>(&words[i]);
            for (int c = 0; c < 6; ++c) acc[c] += h[c + 1];
        }
        uint64_t sum = 0;
        for (int c = 0; c < 6; ++c) sum += acc[c];
        while (sum >= (k << ibits)) ++k;
        if (k < n_frames) {
            frame[k * 7] = i;
            for (int c = 0; c < 6; ++c) frame[k * 7 + 1 + c] = acc[c];
        }
    }
    for (uint64_t f = 1; f < n_frames; ++f)
        if (frame[f * 7] == 0) std::memcpy(&frame[f * 7], &frame[(f - 1) * 7], 56);
}

bool FmdImage::write(const char *fn) const {
    FILE *fp = std::strcmp(fn, "-") ? std::fopen(fn, "wb") : stdout;
    if (!fp) return false;
    const uint32_t a = kAlphabet << 16 | 3;
    const uint64_t zero = 0;
    std::fwrite("RLD\2", 1, 4, fp);
    std::fwrite(&a, 4, 1, fp);
    std::fwrite(&zero, 8, 1, fp);
    std::fwrite(&n_bytes, 8, 1, fp);
    std::fwrite(&n_frames, 8, 1, fp);
    std::fwrite(mcnt + 1, 8, kAlphabet, fp);
    std::fwrite(words.data(), 8, n_bytes / 8, fp);
    std::fwrite(frame.data(), 56, n_frames, fp);
    if (fp != stdout) std::fclose(fp); else std::fflush(fp);
    return true;
}

// ------------------------------------------------------------------------------------ encoder
FmdEncoder::FmdEncoder() {
    w_.assign(1 << 12, 0);
    tail_ = tail_word(0);
}

void FmdEncoder::reserve(uint64_t need) {
    if (need + 4 > w_.size()) {
        uint64_t n = w_.size();
        while (n < need + 4) n <<= 1;
        w_.resize(n, 0);
    }
}

void FmdEncoder::open_block() {
    head_ += kBlockWords;
    reserve(head_ + 2 * kBlockWords);
    uint64_t d[7];
    for (int i = 0; i < 7; ++i) d[i] = tot_[i] - mark_[i], mark_[i] = tot_[i];
    if (d[0] >= 0x8000) {
        uint32_t *h = reinterpret_cast<uint32_t *>(&w_[head_]);
        for (int i = 0; i < 7; ++i) h[i] = (uint32_t)d[i];
        h[0] |= 1u << 31;
        p_ = head_ + kHeaderWords32;
    } else {
        uint16_t *h = reinterpret_cast<uint16_t *>(&w_[head_]);
        for (int i = 0; i < 7; ++i) h[i] = (uint16_t)d[i];
        p_ = head_ + kHeaderWords16;
    }
    tail_ = tail_word(head_);
    room_ = 64;
}

void FmdEncoder::emit(uint64_t len, int sym) {
    const int y = ilog2_32((uint32_t)len), z = ilog2_32((uint32_t)(y + 1));
    int width = 2 * z + 1 + y + 3;       // gamma(y+1) | y mantissa bits | 3-bit symbol
    const uint64_t code = ((((len ^ (1ull << y)) | (uint64_t)(y + 1) << y)) << 3) | (uint64_t)sym;
    if (width >= room_ && p_ == tail_) open_block();
    if (width > room_) {
        width -= room_;
        w_[p_++] |= code >> width;
        room_ = 64 - width;
        w_[p_] = code << room_;
    } else {
        room_ -= width;
        w_[p_] |= code << room_;
    }
    tot_[0] += len; tot_[sym + 1] += len;
}

void FmdEncoder::put(uint64_t len, int sym) {
    if (len == 0) return;
    if (sym != pend_sym_) {
        if (pend_len_) emit(pend_len_, pend_sym_);
        pend_len_ = len; pend_sym_ = sym;
    } else pend_len_ += len;
}

FmdImage FmdEncoder::finish() {
    if (pend_len_) emit(pend_len_, pend_sym_);
    pend_len_ = 0;
    open_block();                         // trailing header-only block (rld.c:230)
    FmdImage e;
    e.n_bytes = p_ * 8;
    w_.resize(p_ + 2);
    w_[p_] = w_[p_ + 1] = 0;
    e.words.swap(w_);
    for (int i = 0; i < 7; ++i) e.mcnt[i] = tot_[i];
    e.finish_counts();
    e.build_frames();
    return e;
}

// ------------------------------------------------------------------------------------ loader
FmdImage *load_fmd(const char *fn) {
    FILE *fp = std::strcmp(fn, "-") ? std::fopen(fn, "rb") : stdin;
    if (!fp) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_fmd_restore] cannot open '%s'\n", fn);
        return nullptr;
    }
    char magic[4];
    if (std::fread(magic, 1, 4, fp) != 4) { if (fp != stdin) std::fclose(fp); return nullptr; }
    std::unique_ptr<FmdImage> e(new FmdImage);
    if (std::memcmp(magic, "RLD\2", 4) == 0) {
        uint32_t a; uint64_t h[3];
        bool ok = std::fread(&a, 4, 1, fp) == 1 && std::fread(h, 8, 3, fp) == 3;
        if (!ok || (a >> 16) != kAlphabet || (a & 0xffff) != 3) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_fmd_restore] '%s': only asize=6, sbits=3 delta-coded .fmd is supported\n", fn);
            if (fp != stdin) std::fclose(fp);
            return nullptr;
        }
        e->n_bytes = h[1]; e->n_frames = h[2];
        ok = std::fread(e->mcnt + 1, 8, kAlphabet, fp) == (size_t)kAlphabet;
        e->finish_counts();
        e->words.assign(e->n_bytes / 8 + 2, 0);
        e->frame.assign(e->n_frames * 7, 0);
        ok = ok && std::fread(e->words.data(), 8, e->n_bytes / 8, fp) == e->n_bytes / 8;
        ok = ok && std::fread(e->frame.data(), 56, e->n_frames, fp) == e->n_frames;
        if (fp != stdin) std::fclose(fp);
        if (!ok) {
            if (fmg_verbose >= 1) std::fprintf(stderr, "[E::fmg_fmd_restore] '%s' is truncated\n", fn);
            return nullptr;
        }
        const uint64_t n_blks = e->n_bytes * 8 / 64 / kBlockWords + 1;
        e->ibits = ilog2_32((uint32_t)(e->mcnt[0] / n_blks)) + 4;
        return e.release();
    }
    // not an RLD file: a raw byte-RLE stream follows the 4 bytes just read (rld.c:295-309)
    FmdEncoder enc;
    std::vector<uint8_t> buf(1 << 16);
    size_t l;
    while ((l = std::fread(buf.data(), 1, buf.size(), fp)) != 0)
        for (size_t i = 0; i < l; ++i)
            if (buf[i] >> 3) enc.put(buf[i] >> 3, buf[i] & 7);
    if (fp != stdin) std::fclose(fp);
    *e = enc.finish();
    return e.release();
}

} // namespace fmg

// ------------------------------------------------------------------------------------ C ABI
using fmg::FmdImage;

struct fmg_fmd_s { FmdImage img; };

extern "C" {

fmg_fmd_t *fmg_fmd_restore(const char *fn) {
    FmdImage *p = fmg::load_fmd(fn);
    if (!p) return nullptr;
    fmg_fmd_t *e = new fmg_fmd_s;
    e->img = std::move(*p);
    delete p;
    return e;
}

fmg_fmd_t *fmg_fmd_from_bwt(int64_t n, const uint8_t *bwt) {
    if (n <= 0 || !bwt) return nullptr;
    fmg::FmdEncoder enc;
    int64_t run = 1; int c = bwt[0];
    for (int64_t i = 1; i < n; ++i) {
        if (bwt[i] != c) { enc.put(run, c); c = bwt[i]; run = 1; }
        else ++run;
    }
    enc.put(run, c);
    fmg_fmd_t *e = new fmg_fmd_s;
    e->img = enc.finish();
    return e;
}

fmg_fmd_t *fmg_fmd_from_rle6(int64_t n, const uint8_t *rle) {
    fmg::FmdEncoder enc;
    for (int64_t i = 0; i < n; ++i)
        if (rle[i] >> 3) enc.put(rle[i] >> 3, rle[i] & 7);
    fmg_fmd_t *e = new fmg_fmd_s;
    e->img = enc.finish();
    return e;
}

fmg_fmd_t *fmg_fmd_from_rld(int asize, int sbits, uint64_t n_bytes, int n_chunks, const uint64_t *const *z,
                            const uint64_t *mcnt, uint64_t n_frames, const uint64_t *frame) {
    if (asize != fmg::kAlphabet || sbits != 3) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] only asize=6, sbits=3 is supported\n", __func__);
        return nullptr;
    }
    fmg_fmd_t *e = new fmg_fmd_s;
    FmdImage &m = e->img;
    m.n_bytes = n_bytes; m.n_frames = n_frames;
    m.words.assign(n_bytes / 8 + 2, 0);
    uint64_t left = n_bytes / 8;
    for (int i = 0; i < n_chunks && left; ++i) {
        const uint64_t take = left < fmg::kChunkWords ? left : fmg::kChunkWords;
        std::memcpy(&m.words[(uint64_t)i * fmg::kChunkWords], z[i], take * 8);
        left -= take;
    }
    for (int c = 1; c <= 6; ++c) m.mcnt[c] = mcnt[c];
    m.finish_counts();
    m.frame.assign(frame, frame + n_frames * 7);
    const uint64_t n_blks = n_bytes * 8 / 64 / fmg::kBlockWords + 1;
    m.ibits = fmg::ilog2_32((uint32_t)(m.mcnt[0] / n_blks)) + 4;
    return e;
}

int fmg_fmd_dump(const fmg_fmd_t *e, const char *fn) {
    if (!e->img.write(fn)) {
        if (fmg_verbose >= 1) std::fprintf(stderr, "[E::%s] cannot write '%s'\n", __func__, fn);
        return -1;
    }
    return 0;
}

void fmg_fmd_destroy(fmg_fmd_t *e) { delete e; }

void fmg_fmd_info(const fmg_fmd_t *e, uint64_t out[17]) {
    for (int i = 0; i < 7; ++i) out[i] = e->img.mcnt[i], out[7 + i] = e->img.cnt[i];
    out[14] = e->img.n_bytes; out[15] = e->img.n_frames; out[16] = e->img.ibits;
}

int64_t fmg_fmd_decode_bwt(const fmg_fmd_t *e, uint8_t *out) {
    int64_t n = 0;
    e->img.for_each_run([&](uint64_t len, int sym) {
        if (out) std::memset(out + n, sym, len);
        n += len;
    });
    return n;
}

void fmg_free(void *p) { std::free(p); }

} // extern "C"
