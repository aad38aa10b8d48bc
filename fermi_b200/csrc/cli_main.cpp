// fermi-b200: fermi's command surface for the hot path (main.c:63-138, cmd.c) on top of libfermi_b200.so.
//   build    cmd.c:378-484   FASTA/Q -> .fmd   (GPU suffix sort; texts < 2^32 symbols)
//   ropebwt  ropebwt.c:47    FASTA/Q -> "RLE\6" byte stream or text BWT (GPU BCR)
//   recode   cmd.c:674-685   RLE\6 / RLD -> RLD .fmd
//   chkbwt   cmd.c:47-120    marginal counts (-p prints the BWT)
//   exact    cmd.c:292-333   SMEMs of every read (GPU), same SQ/EM text as fermi
//   unitig   cmd.c:184-216   MAG records (GPU overlap records + unitig assembly)
//   seqrank  cmd.c:486-505   rank table of fm6_seqsort (GPU fm6_retrieve of every read)
//   kmers    correct.c:305-360  the collect phase of `correct`: "collected N informative and M ambiguous k-mers"
// Host code only parses files and formats text; all index arithmetic happens in the library's CUDA kernels.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <unistd.h>
#include <zlib.h>
#include <sys/time.h>
#include <sys/resource.h>
#include "../../include/fermi_b200.h"

namespace {

const unsigned char nt6_table[128] = {     // seq_nt6_table, seq.c:12-21
    0, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5,
    5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5,
    5, 1, 5, 2, 5, 5, 5, 3, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 4, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5,
    5, 1, 5, 2, 5, 5, 5, 3, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 4, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5};

// minimal FASTA/FASTQ reader over zlib (plain or gzipped), one record at a time
struct SeqReader {
    gzFile fp;
    std::string name, seq, line;
    bool eof = false, have_line = false;
    explicit SeqReader(const char *fn) { fp = std::strcmp(fn, "-") ? gzopen(fn, "rb") : gzdopen(0, "rb"); }
    ~SeqReader() { if (fp) gzclose(fp); }
    bool getline() {
        line.clear();
        char buf[4096];
        for (;;) {
            if (!gzgets(fp, buf, sizeof buf)) { eof = true; return !line.empty(); }
            line += buf;
            if (!line.empty() && line.back() == '\n') { line.pop_back(); if (!line.empty() && line.back() == '\r') line.pop_back(); return true; }
        }
    }
    bool next() {
        if (!fp) return false;
        if (!have_line && !getline()) return false;
        while (line.empty() || (line[0] != '>' && line[0] != '@')) if (!getline()) return false;
        const bool fastq = line[0] == '@';
        const size_t sp = line.find_first_of(" \t");
        name = line.substr(1, sp == std::string::npos ? std::string::npos : sp - 1);
        seq.clear();
        have_line = false;
        while (getline()) {
            if (line.empty()) continue;
            if (line[0] == '>' || (fastq && line[0] == '+') || (!fastq && line[0] == '@')) { have_line = line[0] != '+'; break; }
            seq += line;
        }
        if (fastq && !line.empty() && line[0] == '+') {          // skip the quality string
            size_t got = 0;
            while (got < seq.size() && getline()) got += line.size();
            have_line = false;
        }
        return true;
    }
};

void to_nt6(std::string &s) { for (auto &c : s) c = (unsigned char)c < 128 ? nt6_table[(unsigned char)c] : 5; }
void revcomp6(std::string &s) {
    std::string r(s.rbegin(), s.rend());
    for (auto &c : r) c = (c >= 1 && c <= 4) ? 5 - c : c;
    s.swap(r);
}
bool rc_palindrome(const std::string &s) {                      // cmd.c:458-463, ropebwt.c:25-29
    if (s.size() & 1) return false;
    for (size_t i = 0; i < s.size() / 2; ++i) if (s[i] + s[s.size() - 1 - i] != 5) return false;
    return !s.empty();
}

double realtime() { timeval tp; gettimeofday(&tp, nullptr); return tp.tv_sec + tp.tv_usec * 1e-6; }

int main_build(int argc, char *argv[]) {
    int c, force = 0, no_fr = 1, device = 0;
    const char *out = "-";
    while ((c = getopt(argc, argv, "fo:Od:")) >= 0) {
        if (c == 'f') force = 1; else if (c == 'o') out = optarg; else if (c == 'O') no_fr = 0; else if (c == 'd') device = atoi(optarg);
    }
    if (optind == argc) { std::fprintf(stderr, "Usage: fermi-b200 build [-f] [-O] [-o out.fmd] [-d device] <in.fa>\n"); return 1; }
    if (!force && std::strcmp(out, "-") && access(out, F_OK) == 0) {
        std::fprintf(stderr, "[E::%s] File `%s' exists. Please use `-f' to overwrite.\n", __func__, out); return 1;
    }
    SeqReader rd(argv[optind]);
    if (!rd.fp) { std::fprintf(stderr, "[E::%s] Fail to open the input file.\n", __func__); return 1; }
    std::string text;
    while (rd.next()) {
        to_nt6(rd.seq);
        if (no_fr && rc_palindrome(rd.seq)) rd.seq.pop_back();
        text += rd.seq; text += '\0';
        revcomp6(rd.seq);
        text += rd.seq; text += '\0';
    }
    if (text.empty()) { std::fprintf(stderr, "[E::%s] no sequence in the input\n", __func__); return 1; }
    fmg_fmd_t *e = fmg_build_fmd(device, (int64_t)text.size(), (const uint8_t *)text.data());    // suffix sort + BWT + RLD encoding on the GPU
    const int rc = e ? fmg_fmd_dump(e, out) : 1;
    fmg_fmd_destroy(e);
    return rc != 0;
}

int main_ropebwt(int argc, char *argv[]) {
    int c, bin = 0, fwd = 1, rev = 1, odd = 1, device = 0;
    const char *out = "-";
    int fmd = 0;
    while ((c = getopt(argc, argv, "a:bNtFROo:d:v:r")) >= 0) {         // -a bcr is the only algorithm; -N (cut at N) is always on; -t accepted
        if (c == 'b') bin = 1; else if (c == 'F') fwd = 0; else if (c == 'R') rev = 0; else if (c == 'O') odd = 0;
        else if (c == 'o') out = optarg; else if (c == 'd') device = atoi(optarg); else if (c == 'r') fmd = 1;
    }
    if (optind == argc) {
        std::fprintf(stderr, "Usage: fermi-b200 ropebwt [-b] [-r] [-F] [-R] [-O] [-o out] <in.fq.gz>\n"
                             "       -r writes the RLD-encoded .fmd directly (= `fermi ropebwt -b | fermi recode`, encoded on the GPU)\n");
        return 1;
    }
    SeqReader rd(argv[optind]);
    if (!rd.fp) { std::fprintf(stderr, "[E::%s] Fail to open the input file.\n", __func__); return 1; }
    fmg_bcr_t *b = fmg_bcr_init(device);
    auto insert1 = [&](std::string s) {                          // ropebwt.c:22-45
        if (odd && rc_palindrome(s)) s.pop_back();
        if (s.empty()) return;
        if (fwd) fmg_bcr_append(b, (int)s.size(), (const uint8_t *)s.data());
        if (rev) { revcomp6(s); fmg_bcr_append(b, (int)s.size(), (const uint8_t *)s.data()); }
    };
    while (rd.next()) {
        to_nt6(rd.seq);
        size_t st = 0;
        for (size_t j = 0; j <= rd.seq.size(); ++j)               // cut at ambiguous bases (ropebwt.c:107-116)
            if (j == rd.seq.size() || rd.seq[j] == 5 || rd.seq[j] == 0) { if (j > st) insert1(rd.seq.substr(st, j - st)); st = j + 1; }
    }
    if (fmd) {
        if (fmg_bcr_build(b)) return 1;
        fmg_fmd_t *e = fmg_bcr_fmd(b);
        const int rc = e ? fmg_fmd_dump(e, out) : 1;
        fmg_fmd_destroy(e);
        fmg_bcr_destroy(b);
        return rc != 0;
    }
    if (fmg_bcr_build(b)) return 1;
    FILE *fp = std::strcmp(out, "-") ? std::fopen(out, "wb") : stdout;
    if (!fp) return 1;
    if (bin) {
        uint8_t *rle; int64_t n;
        fmg_bcr_rle(b, &rle, &n);
        std::fwrite("RLE\6", 4, 1, fp);
        std::fwrite(rle, 1, n, fp);
        fmg_free(rle);
    } else {
        std::vector<uint8_t> bwt(fmg_bcr_size(b));
        fmg_bcr_bwt(b, bwt.data());
        for (auto &x : bwt) x = "$ACGTN"[x];
        std::fwrite(bwt.data(), 1, bwt.size(), fp);
        std::fputc('\n', fp);
    }
    if (fp != stdout) std::fclose(fp);
    fmg_bcr_destroy(b);
    return 0;
}

// `fermi merge` (cmd.c:335-373): the indexes are merged left to right on the GPU (gap vector + interleave + RLD encoding)
int main_merge(int argc, char *argv[]) {
    int c, force = 0, device = 0;
    const char *out = "-";
    while ((c = getopt(argc, argv, "fo:t:d:")) >= 0) { if (c == 'f') force = 1; else if (c == 'o') out = optarg; else if (c == 'd') device = atoi(optarg); }
    if (optind + 2 > argc) { std::fprintf(stderr, "Usage: fermi-b200 merge [-f] [-o out.fmd] [-d device] <in0.fmd> <in1.fmd> [...]\n"); return 1; }
    if (!force && std::strcmp(out, "-") != 0) {
        if (FILE *fp = std::fopen(out, "r")) { std::fclose(fp); std::fprintf(stderr, "[E::%s] File `%s' exists. Please use `-f' to overwrite.\n", __func__, out); return 1; }
    }
    fmg_fmd_t *e0 = fmg_fmd_restore(argv[optind]);
    if (!e0) return 1;
    for (int i = optind + 1; i < argc; ++i) {
        fmg_fmd_t *e1 = fmg_fmd_restore(argv[i]);
        fmg_fmd_t *m = e1 ? fmg_merge(e0, e1, device) : nullptr;
        fmg_fmd_destroy(e0);
        if (e1) fmg_fmd_destroy(e1);
        if (!m) return 1;
        e0 = m;
        std::fprintf(stderr, "[M::%s] Merged file `%s' to the existing index.\n", __func__, argv[i]);
    }
    const int rc = fmg_fmd_dump(e0, out);
    fmg_fmd_destroy(e0);
    return rc != 0;
}

int main_recode(int argc, char *argv[]) {
    if (argc < 2) { std::fprintf(stderr, "Usage: fermi-b200 recode <in.rld|in.rle> [out.fmd]\n"); return 1; }
    fmg_fmd_t *e = fmg_fmd_restore(argv[1]);
    if (!e) return 1;
    const int rc = fmg_fmd_dump(e, argc > 2 ? argv[2] : "-");
    fmg_fmd_destroy(e);
    return rc != 0;
}

int main_chkbwt(int argc, char *argv[]) {
    int c, print = 0, check_rank = 0, device = 0;
    while ((c = getopt(argc, argv, "prd:")) >= 0) { if (c == 'p') print = 1; else if (c == 'r') check_rank = 1; else if (c == 'd') device = atoi(optarg); }
    if (optind == argc) { std::fprintf(stderr, "Usage: fermi-b200 chkbwt [-p] [-r] [-d device] <idx.fmd>\n         -r  check the rank function at every position (on the GPU)\n"); return 1; }
    fmg_fmd_t *e = fmg_fmd_restore(argv[optind]);
    if (!e) return 1;
    uint64_t info[17];
    fmg_fmd_info(e, info);
    std::printf("Marginal counts:");
    for (int i = 0; i < 7; ++i) std::printf(" %llu", (unsigned long long)info[i]);
    std::printf("\n");
    if (check_rank) {                                        // cmd.c:90-116: rank1a(k) against the symbols themselves, every k
        fmg_index_t *idx = fmg_index_upload(e, device);
        uint64_t bad = 0, first = 0;
        if (!idx || fmg_check_rank(idx, &bad, &first) != 0) return 1;
        fmg_index_free(idx);
        if (bad) { std::fprintf(stderr, "[E::%s] rank disagrees with the BWT at %llu positions, first at %llu\n", __func__, (unsigned long long)bad, (unsigned long long)first); return 1; }
        std::fprintf(stderr, "[M::%s] Checked the rank function at %llu positions.\n", __func__, (unsigned long long)info[0]);
    }
    if (print) {
        std::vector<uint8_t> bwt(info[0]);
        fmg_fmd_decode_bwt(e, bwt.data());
        for (auto &x : bwt) x = "$ACGTN"[x];
        std::fwrite(bwt.data(), 1, bwt.size(), stdout);
        std::fputc('\n', stdout);
    }
    fmg_fmd_destroy(e);
    return 0;
}

int main_exact(int argc, char *argv[]) {
    int c, self_match = 0, device = 0;
    while ((c = getopt(argc, argv, "Msd:")) >= 0) { if (c == 's') self_match = 1; else if (c == 'd') device = atoi(optarg); }
    if (optind + 2 > argc) { std::fprintf(stderr, "Usage: fermi-b200 exact [-s] [-d device] <idxbase.fmd> <src.fa>\n"); return 1; }
    fmg_fmd_t *e = fmg_fmd_restore(argv[optind]);
    if (!e) return 1;
    uint64_t info[17];
    fmg_fmd_info(e, info);
    fmg_index_t *idx = fmg_index_upload(e, device);
    if (!idx) return 1;
    SeqReader rd(argv[optind + 1]);
    if (!rd.fp) { std::fprintf(stderr, "[E::%s] Fail to open the read file.\n", __func__); return 1; }
    const size_t kBatch = 1 << 20;
    std::vector<std::string> names;
    std::vector<uint8_t> seq;
    std::vector<uint64_t> off{0}, mem_off;
    auto flush = [&]() -> int {
        if (names.empty()) return 0;
        fmg_intv_t *mem = nullptr;
        mem_off.assign(names.size() + 1, 0);
        if (fmg_smem_batch(idx, (int64_t)names.size(), seq.data(), off.data(), self_match, &mem, mem_off.data())) return 1;
        std::string o;
        char buf[128];
        for (size_t i = 0; i < names.size(); ++i) {              // cmd.c:320-327 + fm6_write_smem (smem.c:412-419)
            std::snprintf(buf, sizeof buf, "\t%d\t%d\n", (int)(off[i + 1] - off[i]), (int)(mem_off[i + 1] - mem_off[i]));
            o += "SQ\t"; o += names[i]; o += buf;
            for (uint64_t j = mem_off[i]; j < mem_off[i + 1]; ++j) {
                const fmg_intv_t &a = mem[j];
                std::snprintf(buf, sizeof buf, "EM\t%u\t%u\t%u\t%c%c\n", (unsigned)(a.info >> 32 & 0x3fffffff), (unsigned)(a.info & 0x3fffffff),
                              (unsigned)(a.x[2] > 0xffffffffull ? 0xffffffffu : a.x[2]), "OT"[a.info >> 63], "OT"[a.x[1] < info[1]]);
                o += buf;
            }
            o += "//\n";
        }
        std::fwrite(o.data(), 1, o.size(), stdout);
        fmg_free(mem);
        names.clear(); seq.clear(); off.assign(1, 0);
        return 0;
    };
    while (rd.next()) {
        to_nt6(rd.seq);
        names.push_back(rd.name);
        seq.insert(seq.end(), rd.seq.begin(), rd.seq.end());
        off.push_back(seq.size());
        if (names.size() == kBatch && flush()) return 1;
    }
    if (flush()) return 1;
    fmg_index_free(idx);
    fmg_fmd_destroy(e);
    return 0;
}

int main_unitig(int argc, char *argv[]) {
    int c, min_match = 30, device = 0, max_len = 0;
    while ((c = getopt(argc, argv, "Ml:t:d:L:r:")) >= 0) {
        if (c == 'l') min_match = atoi(optarg); else if (c == 'd') device = atoi(optarg); else if (c == 'L') max_len = atoi(optarg);
        else if (c == 't') setenv("FMG_THREADS", optarg, 1);
        // -r FILE (cmd.c:195): the rank file only lets the reference index its `used` bitmap by row instead of by rank
        // (unitig.c:22-36,282); the unitigs are the same, so the file is accepted and not needed here
    }
    if (optind + 1 > argc) {
        std::fprintf(stderr, "\nUsage:   fermi-b200 unitig [options] <reads.fmd>\n\nOptions: -l INT      min match [%d]\n"
                             "         -t INT      number of host threads of the unitig walk [1]\n         -r FILE     rank file (accepted, not needed)\n"
                             "         -d INT      CUDA device [0]\n\n", min_match);
        return 1;
    }
    fmg_fmd_t *e = fmg_fmd_restore(argv[optind]);
    if (!e) return 1;
    fmg_index_t *idx = fmg_index_upload(e, device);
    if (!idx) return 1;
    const int rc = fmg_unitig(idx, min_match, max_len, "-", nullptr);
    fmg_index_free(idx);
    fmg_fmd_destroy(e);
    return rc != 0;
}

int main_seqsort(int argc, char *argv[]) {              // cmd.c:486-505: the rank table as raw 64-bit words on stdout
    int c, device = 0;
    while ((c = getopt(argc, argv, "t:d:")) >= 0) if (c == 'd') device = atoi(optarg);
    if (optind == argc) { std::fprintf(stderr, "Usage: fermi-b200 seqrank [-d device] <reads.fmd>\n"); return 1; }
    fmg_fmd_t *e = fmg_fmd_restore(argv[optind]);
    fmg_index_t *idx = e ? fmg_index_upload(e, device) : nullptr;
    if (!idx) return 1;
    uint64_t info[17];
    fmg_fmd_info(e, info);
    std::vector<uint64_t> sorted(info[1] ? info[1] : 1);
    int64_t st[3];
    const int rc = fmg_seqsort(idx, sorted.data(), st);
    if (rc == 0) {
        std::fprintf(stderr, "[M::%s] #zeros=%ld, #contained=%ld, #duplicates=%ld\n", "fm6_seqsort", (long)st[0], (long)st[1], (long)st[2]);   // seqsort.c:66
        std::fwrite(sorted.data(), 8, info[1], stdout);
    }
    fmg_index_free(idx);
    fmg_fmd_destroy(e);
    return rc != 0;
}

int main_kmers(int argc, char *argv[]) {
    int c, w = -1, min_occ = 3, device = 0, dump = 0;
    while ((c = getopt(argc, argv, "k:O:d:p")) >= 0) {
        if (c == 'k') w = atoi(optarg); else if (c == 'O') min_occ = atoi(optarg); else if (c == 'd') device = atoi(optarg); else if (c == 'p') dump = 1;
    }
    if (optind == argc) { std::fprintf(stderr, "Usage: fermi-b200 kmers [-k kmer] [-O minOcc] [-p] <reads.fmd>\n"); return 1; }
    fmg_fmd_t *e = fmg_fmd_restore(argv[optind]);
    fmg_index_t *idx = e ? fmg_index_upload(e, device) : nullptr;
    if (!idx) return 1;
    uint64_t *tri, n; int64_t cnt[2];
    const double t0 = realtime();
    if (fmg_ec_collect(idx, w, min_occ, &tri, &n, cnt)) return 1;
    std::fprintf(stderr, "[M::%s] collected %ld informative and %ld ambiguous k-mers in %.3f sec wall clock\n", __func__, (long)cnt[1],
                 (long)(cnt[0] - cnt[1]), realtime() - t0);                       // correct.c:359-360
    if (dump) for (uint64_t i = 0; i < n; ++i) std::printf("%llu\t%llu\t%llu\n", (unsigned long long)(tri[i] >> 40),
                                                            (unsigned long long)(tri[i] >> 8 & 0xffffffffull), (unsigned long long)(tri[i] & 0xff));
    fmg_free(tri);
    fmg_index_free(idx);
    fmg_fmd_destroy(e);
    return 0;
}

int usage() {
    std::fprintf(stderr, "\nProgram: fermi-b200 (FMD-index hot path of fermi on NVIDIA B200)\nVersion: %s\n\n", FMG_VERSION);
    std::fprintf(stderr, "Usage:   fermi-b200 <command> [arguments]\n\n");
    std::fprintf(stderr, "Command: build      generate the FMD-index (GPU suffix sort)\n");
    std::fprintf(stderr, "         ropebwt    BWT of a read set by BCR on the GPU\n");
    std::fprintf(stderr, "         recode     convert RLE\\6 / RLD to RLD\n");
    std::fprintf(stderr, "         merge      merge FMD-indexes (gap vector on the GPU)\n");
    std::fprintf(stderr, "         chkbwt     marginal counts / print the BWT / check the rank function\n");
    std::fprintf(stderr, "         exact      find supermaximal exact matches\n");
    std::fprintf(stderr, "         unitig     construct unitigs\n");
    std::fprintf(stderr, "         seqrank    compute the rank of sequences\n");
    std::fprintf(stderr, "         kmers      k-mer collection of `correct`\n\n");
    return 1;
}

} // namespace

int main(int argc, char *argv[]) {                    // main.c:63-138
    if (argc < 2) return usage();
    const double t0 = realtime();
    int ret;
    const std::string cmd = argv[1];
    if (cmd == "build") ret = main_build(argc - 1, argv + 1);
    else if (cmd == "ropebwt") ret = main_ropebwt(argc - 1, argv + 1);
    else if (cmd == "recode") ret = main_recode(argc - 1, argv + 1);
    else if (cmd == "merge") ret = main_merge(argc - 1, argv + 1);
    else if (cmd == "chkbwt") ret = main_chkbwt(argc - 1, argv + 1);
    else if (cmd == "exact") ret = main_exact(argc - 1, argv + 1);
    else if (cmd == "unitig") ret = main_unitig(argc - 1, argv + 1);
    else if (cmd == "seqsort" || cmd == "seqrank") ret = main_seqsort(argc - 1, argv + 1);      // main.c:108-109
    else if (cmd == "kmers") ret = main_kmers(argc - 1, argv + 1);
    else { std::fprintf(stderr, "[E::%s] unrecognized command '%s'\n", __func__, argv[1]); return 1; }
    if (ret == 0 && fmg_verbose >= 3) {
        rusage r; getrusage(RUSAGE_SELF, &r);
        std::fprintf(stderr, "[M::%s] Version: %s\n[M::%s] CMD:", __func__, FMG_VERSION, __func__);
        for (int i = 0; i < argc; ++i) std::fprintf(stderr, " %s", argv[i]);
        std::fprintf(stderr, "\n[M::%s] Real time: %.3f sec; CPU: %.3f sec\n", __func__, realtime() - t0,
                     r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec));
    }
    return ret;
}
