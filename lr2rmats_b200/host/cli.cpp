// cli.cpp -- drop-in `lr2rmats` command line above the C ABI.
//
// Same subcommands, option letters, long-option names (quirks included: `-M` takes an argument, `--source` maps to 's',
// bam2gtf's long names are exon-min / intron-len -- SURVEY.md App. A.5/A.9/D.1), defaults, output files and exit codes as
// main.c:37-49, bam_filter.c:98-164, bam2gtf.c:120-161, update_gtf.c:995-1117, unique_gtf.c:86-158.  The per-alignment
// work itself is done by the engine (the CUDA library in the product binary).
#include <unistd.h>
#include <zlib.h>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <getopt.h>
#include <string>
#include <unordered_map>
#include <vector>
#include "lrb_host.h"

namespace lrb {

static const char *PROG = "lr2rmats";

static void logf(const char *func, const char *msg)
{
    time_t raw; time(&raw); char buf[80]; strftime(buf, 80, "%m-%d-%Y %X", localtime(&raw));
    fprintf(stderr, "=== %s === [%s] %s", buf, func, msg);
}
// _exit, not exit: the CUDA context may still be under construction on the helper thread (main.cpp), and exit() would run the
// runtime's static destructors underneath it
[[noreturn]] static void fatal(const char *func, const std::string &msg) { fprintf(stderr, "[%s] %s\n", func, msg.c_str()); fflush(NULL); _exit(EXIT_FAILURE); }
static void engine_fail(Engine &e, const char *func, int rc)
{
    const char *m = e.error ? e.error(e.self) : nullptr;
    fatal(func, std::string("device path failed (code ") + std::to_string(rc) + "): " + (m ? m : "?"));
}

static int usage()
{
    fprintf(stderr, "\n");
    fprintf(stderr, "Program: %s (%s)\n", PROG, "Long read to rMATS");
    fprintf(stderr, "Version: %s, Date: %s\n", "0.1", "2018-02-09");
    fprintf(stderr, "Contact: %s\n", "yangaoucla@gmail.com");
    fprintf(stderr, "Usage:   %s <command> [options]\n\n", PROG);
    fprintf(stderr, "Commands: \n");
    fprintf(stderr, "         filter       filter out alignment records with low confidence\n");
    fprintf(stderr, "         fusion       generate candidate gene-fusion transcripts\n");
    fprintf(stderr, "         update-gtf   generate new GTF file based on BAM/SAM and existing GTF file\n");
    fprintf(stderr, "         unique-gtf   generate GTF file that only contain unique transcript based on BAM/SAM or GTF file\n");
    fprintf(stderr, "         bam2gtf      generate transcript and exon information based on BAM/SAM file\n");
    fprintf(stderr, "         bam2sj       generate splice-junction information based on BAM/SAM file\n");
    fprintf(stderr, "\n");
    return 1;
}

// ------------------------------------------------------------------------------------------------ filter
static int filter_usage()
{
    fprintf(stderr, "\n");
    fprintf(stderr, "Usage:   %s filter [option] <in.bam/sam> | samtools sort > out.sort.bam\n\n", PROG);
    fprintf(stderr, "Options:\n");
    fprintf(stderr, "         -v --coverage   [FLOAT]    minimum fraction of aligned bases. [%.2f]\n", 0.67);
    fprintf(stderr, "         -q --map-qual   [FLOAT]    minimum fraction of identically aligned bases. [%.2f]\n", 0.75);
    fprintf(stderr, "         -s --sec-rat    [FLOAT]    maximum ratio of second best and best score to retain the best\n");
    fprintf(stderr, "                                    alignment, or no alignments will be retained. [%.2f]\n", 0.98);
    fprintf(stderr, "         -i --intron     [INT]      minimum number of intron indicated by the alignment. [%d]\n", 0);
    fprintf(stderr, "         -r --remove-gtf [STR]      remove all the alignment record that overlap with transcript in this GTF file. [NONE]\n");
    fprintf(stderr, "\n");
    return 1;
}

static int cmd_filter(int argc, char **argv, Engine &eng)
{
    static const struct option lo[] = {{"coverage", 1, NULL, 'v'}, {"map-quality", 1, NULL, 'q'}, {"sec-rat", 1, NULL, 's'},
                                       {"intron", 1, NULL, 'i'}, {"remove-gtf", 1, NULL, 'r'}, {0, 0, 0, 0}};
    lrb_filter_params fp = {(float)0.67, (float)0.75, (float)0.98, 0};
    std::string rm_fn; int c;
    while ((c = getopt_long(argc, argv, "v:q:s:i:r:", lo, NULL)) >= 0) {
        switch (c) {
        case 'v': fp.cov_rate = (float)atof(optarg); break;
        case 'q': fp.map_qual = (float)atof(optarg); break;
        case 's': fp.sec_rat = (float)atof(optarg); break;
        case 'i': fp.min_intron_n = atoi(optarg); break;
        case 'r': rm_fn = optarg; break;
        default: return filter_usage();
        }
    }
    if (argc - optind != 1) return filter_usage();
    Header h; Records rec; rec.keep_raw = true; std::string err;
    IoTrace tr;
    if (!read_alignments(argv[optind], h, rec, err)) fatal("bam_filter", err);
    tr.lap("alignments");
    Anno rm;
    if (!rm_fn.empty()) {
        logf("read_anno_trans", ("reading transcript annotation from " + rm_fn + " ...\n").c_str());
        if (!read_gtf(rm_fn, h, rm, false, err)) fatal("read_anno_trans", err);
        logf("read_anno_trans", ("reading transcript annotation from " + rm_fn + " done.\n").c_str());
    }
    lrb_anno rmv = rm.view();
    int rc = eng.set_tables(eng.self, nullptr, rm_fn.empty() ? nullptr : &rmv, nullptr);
    if (rc) engine_fail(eng, "bam_filter", rc);
    tr.lap("tables upload");
    lrb_batch b = rec.view(); lrb_filter_result res;
    for (int attempt = 0;; ++attempt) {
        rc = eng.filter(eng.self, &b, &fp, &res);
        if (rc) engine_fail(eng, "bam_filter", rc);
        // The device forms the qname runs (bam_filter.c:129-159, strcmp on the passing subsequence) on 64-bit hashes of the names.  Every
        // pair it took for equal is checked against the names here; two different names with one hash (2^-64 per pair) get fresh
        // hashes and the stage runs again.
        bool clash = false; int64_t prev = -1;
        for (int64_t i = 0; i < res.n; ++i) {
            if (!res.pass[i]) continue;
            if (prev >= 0 && rec.qhash[(size_t)i] == rec.qhash[(size_t)prev] && strcmp(rec.qname((size_t)i), rec.qname((size_t)prev)) != 0) { clash = true; break; }
            prev = i;
        }
        if (!clash) break;
        if (attempt == 3) fatal("bam_filter", "query-name hash collisions persist");
        for (size_t i = 0; i < rec.n(); ++i) {        // another hash function of the NAME (a function of the old hash would collide again)
            uint64_t hsh = 0xCBF29CE484222325ull ^ ((uint64_t)(attempt + 1) * 0xD1B54A32D192ED03ull);
            for (const char *q = rec.qname(i); *q; ++q) { hsh ^= (uint8_t)*q; hsh *= 0x100000001B3ull; hsh ^= hsh >> 31; }
            rec.qhash[i] = hsh;
        }
        b = rec.view();
    }
    tr.lap("engine");
    if (!write_bam(stdout, h, rec, res.keep_idx, res.n_keep, err)) fatal("bam_filter", err);
    tr.lap("emit");
    logf("bam_filter", ("Filtered alignments: " + std::to_string(res.n_keep) + "\n").c_str());
    return 0;
}

// ----------------------------------------------------------------------------------------------- bam2gtf
static int bam2gtf_usage()
{
    fprintf(stderr, "\n");
    fprintf(stderr, "Usage:   %s bam2gtf [option] <in.bam> > out.gtf\n\n", PROG);
    fprintf(stderr, "Options:\n\n");
    fprintf(stderr, "         -e --min-exon    [INT]    minimum length of internal exon. [%d]\n", 3);
    fprintf(stderr, "         -i --min-intron  [INT]    minimum length of intron. [%d]\n", 3);
    fprintf(stderr, "         -t --max-delet   [INT]    maximum length of deletion, longer deletion will be considered as intron. [%d]\n", 50);
    fprintf(stderr, "         -s --source      [STR]    source field in GTF, program, database or project name. [%s]\n", PROG);
    fprintf(stderr, "\n");
    return 1;
}

static int cmd_bam2gtf(int argc, char **argv, Engine &eng)
{
    static const struct option lo[] = {{"exon-min", 1, NULL, 'e'}, {"intron-len", 1, NULL, 'i'}, {"source", 1, NULL, 's'}, {0, 0, 0, 0}};
    lrb_exon_params ep = {3, 3, 50}; std::string src = PROG; int c;
    while ((c = getopt_long(argc, argv, "s:e:i:t:", lo, NULL)) >= 0) {
        switch (c) {
        case 'e': ep.min_exon = atoi(optarg); break;
        case 'i': ep.min_intron = atoi(optarg); break;
        case 't': ep.max_delet = atoi(optarg); break;
        case 's': src = optarg; break;
        default: fprintf(stderr, "Error: unknown option: %s.\n", optarg); return bam2gtf_usage();
        }
    }
    if (argc - optind != 1) return bam2gtf_usage();
    Header h; Records rec; std::string err;
    if (!read_alignments(argv[optind], h, rec, err)) fatal("bam2gtf", err);
    ChrNames cn; cn.seed(h);
    lrb_batch b = rec.view(); lrb_exon_result res;
    int rc = eng.bam2gtf(eng.self, &b, &ep, &res);
    if (rc) engine_fail(eng, "bam2gtf", rc);
    emit_bam2gtf(stdout, res, rec, cn, src.c_str());
    return 0;
}

// ------------------------------------------------------------------------------------------------ bam2sj
// parse_bam.c:44-69 (usage), :987-1060 (bam2sj): option letters, long names and the output layout are the drop-in contract
static int bam2sj_usage()
{
    fprintf(stderr, "\n");
    fprintf(stderr, "Usage:   %s bam2sj [option] <in.bam> > out.sj\n\n", PROG);
    fprintf(stderr, "Note:    in.bam should be sorted in advance\n\n");
    fprintf(stderr, "Input Options:\n\n");
    fprintf(stderr, "         -G --gtf-anno    [STR]    GTF annotation file, indicating known splice-junctions. \n");
    fprintf(stderr, "         -g --genome-file [STR]    genome.fa. Use genome sequence to classify intron-motif. \n");
    fprintf(stderr, "                                   If no genome file is give, intron-motif will be set as 0\n");
    fprintf(stderr, "                                   (non-canonical) [None]\n");
    fprintf(stderr, "\nFilter Options:\n\n");
    fprintf(stderr, "         -p --prop-pair            set -p to force to filter out reads mapped in improper pair. [False]\n");
    fprintf(stderr, "         -a --anchor-len  [INT,INT,INT,INT,INT]\n");
    fprintf(stderr, "                                   minimum anchor length for junction read, [annotated, non-canonical,\n");
    fprintf(stderr, "                                    GT/AG, GC/AG, AT/AC]. [%d,%d,%d,%d,%d]\n", 1, 30, 12, 12, 12);
    fprintf(stderr, "         -U --uniq-map    [INT,INT,INT,INT,INT]\n");
    fprintf(stderr, "                                   minimum uniq-map read count for junction read, [annotated,\n");
    fprintf(stderr, "                                   non-canonical, GT/AG, GC/AG, AT/AC]. [%d,%d,%d,%d,%d]\n", 0, 3, 1, 1, 1);
    fprintf(stderr, "         -A --all-map     [INT,INT,INT,INT,INT]\n");
    fprintf(stderr, "                                   minimum total uniq-map and multi-map read count for junction\n");
    fprintf(stderr, "                                   read, [annotated, non-canonical, GT/AG, GC/AG, AT/AC].\n");
    fprintf(stderr, "                                   [%d,%d,%d,%d,%d]\n", 0, 3, 1, 1, 1);
    fprintf(stderr, "         -i --intron-len  [INT]    minimum intron length for junction read. [%d]\n", 3);
    fprintf(stderr, "\n");
    return 1;
}
// kseq_load_genome (parse_bam.c:382-400): sequences in file order (bam2sj indexes them by tid), plain or gzip
static bool load_genome(const char *fn, std::vector<std::string> &seqs)
{
    gzFile fp = gzopen(fn, "r");
    if (!fp) return false;
    std::vector<char> buf(1 << 22); std::string line; int n; bool in_seq = false;
    auto feed = [&](const std::string &l) {
        if (!l.empty() && l[0] == '>') { seqs.emplace_back(); in_seq = true; }
        else if (in_seq) { for (char ch : l) if (!isspace((unsigned char)ch)) seqs.back().push_back(ch); }
    };
    while ((n = gzread(fp, buf.data(), (unsigned)buf.size())) > 0) {
        const char *p = buf.data(), *e = p + n;
        while (p < e) { const char *q = (const char *)memchr(p, '\n', (size_t)(e - p)); if (!q) { line.append(p, e); break; } line.append(p, q); feed(line); line.clear(); p = q + 1; }
    }
    if (!line.empty()) feed(line);
    gzclose(fp);
    return true;
}
static int five_ints(const char *arg) { char *p; strtol(arg, &p, 10); for (int k = 0; k < 4; ++k) { if (*p == 0) return -1; strtol(p + 1, &p, 10); } return 0; }
static int cmd_bam2sj(int argc, char **argv, Engine &eng)
{
    static const struct option lo[] = {{"proper-pair", 1, NULL, 'p'}, {"gtf-anno", 1, NULL, 'G'}, {"genome-file", 1, NULL, 'g'}, {"anchor-len", 1, NULL, 'a'},
                                       {"uniq-map", 1, NULL, 'U'}, {"all-map", 1, NULL, 'A'}, {"intron-len", 1, NULL, 'i'}, {0, 0, 0, 0}};
    lrb_sj_params sp = {3, 1}; std::string ref_fn; int c;
    while ((c = getopt_long(argc, argv, "G:g:pa:i:A:U:", lo, NULL)) >= 0) {
        switch (c) {
        case 'g': ref_fn = optarg; break;
        case 'p': sp.pair_only = 1; break;                              // PAIR_T again: no option reaches single-end mode (parse_bam.c:997)
        case 'a': case 'U': case 'A': if (five_ints(optarg)) return bam2sj_usage(); break;      // parsed, never read by bam2sj_core
        case 'i': sp.min_intron = atoi(optarg); break;
        default: fprintf(stderr, "Error: unknown option: %s.\n", optarg); return bam2sj_usage();
        }
    }
    if (argc - optind != 1) return bam2sj_usage();
    std::vector<std::string> genome;
    if (!ref_fn.empty()) {
        if (!eng.bam2sj) fatal("bam2sj", "this engine has no bam2sj");
        logf("kseq_load_genome", "loading genome fasta file ...\n");
        if (!load_genome(ref_fn.c_str(), genome)) fatal("bam2sj", "Can not open genome file. " + ref_fn);
        logf("kseq_load_genome", "loading genome fasta file done!\n");
    }
    Header h; Records rec; std::string err;
    if (!read_alignments(argv[optind], h, rec, err)) fatal("bam2sj", err);
    logf("bam2sj_core", "generating splice-junction with BAM file ...\n");
    std::vector<uint8_t> uniq(rec.n());
    for (size_t i = 0; i < rec.n(); ++i) {
        if (rec.flag[i] & 4) continue;
        if (rec.nh[i] == 0) fprintf(stderr, "No \"NH\" tag.\n");       // bam_is_uniq_NH, parse_bam.c:239-247: once per mapped record, before the pair test
        uniq[i] = rec.nh[i] == 1;
    }
    lrb_batch b = rec.view(); lrb_sj res;
    int rc = eng.bam2sj ? eng.bam2sj(eng.self, &b, uniq.data(), &sp, &res) : LRB_E_ARG;
    if (rc) engine_fail(eng, "bam2sj", rc);
    logf("bam2sj_core", "generating splice-junction with BAM file done!\n");
    // print_sj (parse_bam.c:974-985); strand / motif from the genome (intr_deri_str :319-337), 0 / 0 without -g
    static const char motifs[6][5] = {"GTAG", "CTAC", "GCAG", "CTGC", "ATAC", "GTAT"};
    static const int motif_strand[6] = {1, 2, 1, 2, 1, 2};
    std::string out;
    out += "###STRAND 0:undefined, 1:+, 2:-\n###ANNO 0:novel, 1:annotated\n###MOTIF 0:non-canonical, 1:GT/AG, 2:CT/AC, 3:GC/AG, 4:CT/GC, 5:AT/AC, 6:GT/AT\n#CHR\tSTART\tEND\tSTRAND\tANNO\tUNIQ_C\tMULTI_C\tMOTIF\n";
    char line[256];
    for (int64_t k = 0; k < res.n; ++k) {
        int strand = 0, motif = 0;
        const int tid = res.tid[k], don = res.don[k], acc = res.acc[k];
        if (!genome.empty()) {
            if (tid >= (int)genome.size()) fatal("intr_deri_str", "unknown tid: " + std::to_string(tid) + "\n");
            const std::string &g = genome[(size_t)tid];
            auto at = [&](long pos) { return pos >= 0 && pos < (long)g.size() ? (char)toupper((unsigned char)g[(size_t)pos]) : '\0'; };
            const char in[5] = {at(don - 1), at(don), at(acc - 2), at(acc - 1), 0};
            for (int m = 0; m < 6; ++m) if (strcmp(in, motifs[m]) == 0) { motif = m + 1; strand = motif_strand[m]; break; }
        }
        const char *name = tid >= 0 && tid < (int)h.names.size() ? h.names[(size_t)tid].c_str() : "*";
        out.append(line, (size_t)snprintf(line, sizeof line, "%s\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", name, don, acc, strand, 1, res.uniq_c[k], res.multi_c[k], motif));
    }
    fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}

// ---------------------------------------------------------------------------------------------- sort-gtf
// The last step of the pipeline sorts the concatenated GTF with src/sort_gtf.sh (Snakefile:192): an awk pass tags every `transcript` /
// `exon` line with (chromosome rank, start, end of the last transcript line, line number), `sort -n -k1 -n -k2 -n -k3 -n -k4` orders
// the tagged lines and a second awk pass prints the first nine tab-separated columns.  Here the tagging is one pass on the host (it
// carries state from line to line), the sort runs on the device (a stable radix sort of the three keys; the line number is the input
// order), and the lines are written through the permutation.  `lr2rmats-b200 sort-gtf in.unsort.gtf out.sort.gtf` replaces the script.
static int sort_gtf_usage() { fprintf(stdout, "Usage: %s sort-gtf in.unsort.gtf out.sort.gtf\n       sort GTF file based on 'trans' lines\n", PROG); return 0; }
static int cmd_sort_gtf(int argc, char **argv, Engine &eng)
{
    if (argc != 3) return sort_gtf_usage();                            // sort_gtf.sh:2-6: usage on stdout, exit status 0
    if (!eng.sort3) fatal("sort-gtf", "this engine has no sort");
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) fatal("sort-gtf", std::string("Cannot open \"") + argv[1] + "\"");
    std::string text; { char buf[1 << 16]; size_t k; while ((k = fread(buf, 1, sizeof buf, fp)) > 0) text.append(buf, k); } fclose(fp);
    struct Line { size_t off, len; };
    std::vector<Line> lines; std::vector<uint32_t> k0, k1, k2;
    std::unordered_map<std::string, uint32_t> chrom;
    { static const char *fixed[] = {"chr1", "chr2", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16",
                                    "chr17", "chr18", "chr19", "chr20", "chr21", "chr22", "chrX", "chrY", "chrM"};
      for (uint32_t i = 0; i < 25; ++i) chrom[fixed[i]] = i + 1; }
    uint32_t chr = 0, chr_m = 25; long long start = 0, end = 0; bool end_set = false; uint64_t nr = 0;
    auto is_blank = [](char ch) { return ch == ' ' || ch == '\t'; };
    auto lead_int = [](const char *p, const char *e) { long long v = 0; bool neg = false; while (p < e && (*p == ' ' || *p == '\t')) ++p; if (p < e && *p == '-') { neg = true; ++p; }
                                                        while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); if (v > (1ll << 40)) break; ++p; } return neg ? -v : v; };
    for (size_t p = 0; p < text.size();) {
        size_t q = text.find('\n', p); if (q == std::string::npos) q = text.size();
        ++nr;
        const char *b = text.data() + p, *e = text.data() + q;
        if (!(b < e && *b == '#')) {
            // awk's default field splitting: runs of blanks separate fields, leading blanks are skipped
            const char *f[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *fe[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
            const char *c = b; int nf = 0;
            while (nf < 5) { while (c < e && is_blank(*c)) ++c; if (c >= e) break; f[nf] = c; while (c < e && !is_blank(*c)) ++c; fe[nf] = c; ++nf; }
            if (nf >= 3) {
                const std::string t3(f[2], fe[2]);
                if (t3.find("transcript") != std::string::npos || t3 == "exon") {
                    if (t3 == "transcript") {
                        const std::string c1(f[0], fe[0]);
                        auto it = chrom.find(c1);
                        if (it == chrom.end()) it = chrom.emplace(c1, ++chr_m).first;
                        chr = it->second; start = nf > 3 ? lead_int(f[3], fe[3]) : 0; end = nf > 4 ? lead_int(f[4], fe[4]) : 0; end_set = true;
                    }
                    if (start < 0 || end < 0 || start > 0xffffffffll || end > 0xffffffffll || nr > 0xffffffffull) fatal("sort-gtf", "coordinate or line number beyond 2^32");
                    // before the first transcript line awk's `end` is still empty: `sort` then sees the line number as its third field
                    lines.push_back({p, q - p}); k0.push_back(chr); k1.push_back((uint32_t)start); k2.push_back(end_set ? (uint32_t)end : (uint32_t)nr);
                }
            }
        }
        p = q + 1;
    }
    const uint32_t *perm = nullptr;
    int rc = eng.sort3(eng.self, k0.data(), k1.data(), k2.data(), (int64_t)lines.size(), &perm);
    if (rc) engine_fail(eng, "sort-gtf", rc);
    FILE *out = fopen(argv[2], "wb");
    if (!out) fatal("sort-gtf", std::string("Cannot open \"") + argv[2] + "\"");
    std::string o; o.reserve(text.size() + lines.size() * 8);
    for (size_t i = 0; i < lines.size(); ++i) {
        const Line &l = lines[perm[i]];
        // the second awk pass (FS = tab) prints nine fields whatever the line holds
        const char *b = text.data() + l.off, *e = b + l.len; int col = 0;
        while (col < 9) {
            const char *t = (const char *)memchr(b, '\t', (size_t)(e - b)); if (!t) t = e;
            o.append(b, t); ++col; if (col < 9) o.push_back('\t');
            b = t < e ? t + 1 : e;
        }
        o.push_back('\n');
    }
    fwrite(o.data(), 1, o.size(), out); fclose(out);
    return 0;
}

// -------------------------------------------------------------------------------------------- update-gtf
static int update_usage()
{
    fprintf(stderr, "\n");
    fprintf(stderr, "Usage:   %s update-gtf [option] <in.bam/in.gtf> <old.gtf> > new.gtf\n\n", PROG);
    fprintf(stderr, "Notice:  the BAM and GTF files should be sorted in advance.\n\n");
    fprintf(stderr, "Input options:\n\n");
    fprintf(stderr, "         -m --input-mode   [STR]    format of input file <in.bam/in.gtf>, BAM file(b) or GTF file(g). [b]\n");
    fprintf(stderr, "         -b --bam          [STR]    for GTF input <in.gtf>, BAM file is needed to obtain BAM header information. [NULL]\n");
    fprintf(stderr, "         -j --sj           [STR]    junction information file output by STAR(*.out.tab). [NULL]\n");
    fprintf(stderr, "\n");
    fprintf(stderr, "Function options:\n\n");
    fprintf(stderr, "         -c --force-strand         force to match strand when merging transcripts. [False]\n");
    fprintf(stderr, "         -e --min-exon     [INT]    minimum length of internal exon. [%d]\n", 3);
    fprintf(stderr, "         -i --min-intron   [INT]    minimum length of intron. [%d]\n", 3);
    fprintf(stderr, "         -t --max-delet    [INT]    maximum length of deletion, longer deletion will be considered as intron. [%d]\n", 50);
    fprintf(stderr, "         -d --distance     [INT]    consider same if distance between two splice site is not bigger than d. [%d]\n", 0);
    fprintf(stderr, "         -D --DISTANCE     [INT]    consider same if distance between two start/end site is not bigger than D. [%d]\n", 0x7fffffff);
    fprintf(stderr, "         -f --frac         [INT]    consider same if overlapping between two single-exon transcript is bigger than f. [%.2f]\n", 0.80);
    fprintf(stderr, "         -s --split-trans           split read on unreliable junctions. [False]\n");
    fprintf(stderr, "         -M --use-multi             use junction information of multi-mapped read. [False]\n");
    fprintf(stderr, "         -J --min-junc-cnt [INT]    minimum short-read junction count of novel junction. [%d]\n", 1);
    fprintf(stderr, "         -l --full-length  [INT]    level of strict criterion for considering full-length transcript. \n");
    fprintf(stderr, "                                    (1->5, most strict->most relaxed) [%d]\n", 5);
    fprintf(stderr, "\n");
    fprintf(stderr, "Output options:\n\n");
    fprintf(stderr, "         -o --output       [STR]    updated GTF file. [stdout]\n");
    fprintf(stderr, "         -n --min-output            only keep the minimal set of novel transcripts in the updated GTF file. [False]\n");
    fprintf(stderr, "         -E --exon-bed     [STR]    updated novel exon file in bed format. [NULL]\n");
    fprintf(stderr, "         -a --bam-gtf      [STR]    bam-derived transcript GTF file. [NULL]\n");
    fprintf(stderr, "         -A --bam-detial   [STR]    detailed information of each bam-derived transcript. [NULL]\n");
    fprintf(stderr, "         -k --known-gtf    [STR]    bam-derived known transcript GTF file. [NULL]\n");
    fprintf(stderr, "         -v --novel-gtf    [STR]    bam-derived novel transcript GTF file. [NULL]\n");
    fprintf(stderr, "         -u --unrecog      [STR]    bam-derived unrecognized transcript GTF file. [NULL]\n");
    fprintf(stderr, "         -y --summary      [STR]    Staticstic summary of bam-derived transcript. [NULL]\n");
    fprintf(stderr, "         -S --source       [STR]    \'source\' field in GTF: program, database or project name. [%s]\n", PROG);
    fprintf(stderr, "\n");
    return 1;
}

static void default_update_params(lrb_update_params &up)
{
    up.min_sj_cnt = 1; up.ss_dis = 0; up.end_dis = 0x7fffffff; up.full_level = 5; up.split_trans = 0; up.use_multi = 0;
    up.force_strand = 0; up.single_exon_ovlp_frac = (float)0.80; up.want_summary = 0;
}

static int cmd_update(int argc, char **argv, Engine &eng)
{
    static const struct option lo[] = {
        {"input-mode", 1, NULL, 'm'}, {"bam", 1, NULL, 'b'}, {"sj", 1, NULL, 'j'}, {"force-strand", 0, NULL, 'c'},
        {"min-exon", 1, NULL, 'e'}, {"min-intron", 1, NULL, 'i'}, {"distance", 1, NULL, 'd'}, {"DISTANCE", 1, NULL, 'D'},
        {"frac", 1, NULL, 'f'}, {"full-gtf", 1, NULL, 'l'}, {"use-multi", 0, NULL, 'M'}, {"min_sj_cnt", 1, NULL, 'J'},
        {"output", 1, NULL, 'o'}, {"bam-gtf", 1, NULL, 'a'}, {"known-gtf", 1, NULL, 'k'}, {"novel-gtf", 1, NULL, 'v'},
        {"unrecog", 1, NULL, 'u'}, {"source", 1, NULL, 's'}, {0, 0, 0, 0}};
    lrb_update_params up; default_update_params(up);
    lrb_exon_params ep = {3, 3, 50};
    int input_mode = 0; std::string gtf_bam, sj_fn, src = PROG;
    FILE *out_fp = stdout, *bed_fp = NULL, *bam_gtf_fp = NULL, *detail_fp = NULL, *known_fp = NULL, *novel_fp = NULL, *unrecog_fp = NULL, *summary_fp = NULL;
    int c;
    while ((c = getopt_long(argc, argv, "m:b:j:J:M:e:i:t:sd:D:f:cl:o:nE:a:A:k:v:u:y:S:", lo, NULL)) >= 0) {
        switch (c) {
        case 'm': if (optarg[0] == 'b') input_mode = 0; else if (optarg[0] == 'g') input_mode = 1; else return update_usage(); break;
        case 'b': gtf_bam = optarg; { FILE *t = fopen(optarg, "rb"); if (!t) fatal("update_gtf", std::string("Cannot open \"") + optarg + "\"\n"); fclose(t); } break;
        case 'j': sj_fn = optarg; { FILE *t = fopen(optarg, "r"); if (!t) fatal("update_gtf", std::string("Can not open splice-junction file \"") + optarg + "\"\n"); fclose(t); } break;
        case 'e': ep.min_exon = atoi(optarg); break;
        case 'i': ep.min_intron = atoi(optarg); break;
        case 't': ep.max_delet = atoi(optarg); break;
        case 'd': up.ss_dis = atoi(optarg); break;
        case 'D': up.end_dis = atoi(optarg); break;
        case 'f': up.single_exon_ovlp_frac = (float)atof(optarg); break;
        case 'c': up.force_strand = 1; break;
        case 's': up.split_trans = 1; break;
        case 'l': up.full_level = atoi(optarg); break;
        case 'M': up.use_multi = 1; break;
        case 'J': up.min_sj_cnt = atoi(optarg); break;
        case 'o': out_fp = fopen(optarg, "w"); break;
        case 'n': break;
        case 'E': bed_fp = fopen(optarg, "w"); break;
        case 'a': bam_gtf_fp = fopen(optarg, "w"); break;
        case 'A': detail_fp = fopen(optarg, "w"); break;
        case 'k': known_fp = fopen(optarg, "w"); break;
        case 'v': novel_fp = fopen(optarg, "w"); break;
        case 'u': unrecog_fp = fopen(optarg, "w"); break;
        case 'y': summary_fp = fopen(optarg, "w"); break;
        case 'S': src = optarg; break;
        default: fprintf(stderr, "Error: unknown option: %s.\n", optarg); return update_usage();
        }
    }
    if (argc - optind != 2) return update_usage();
    up.want_summary = (summary_fp || bed_fp) ? 1 : 0;

    Header h; Records rec; Anno chains, anno; ChrNames cn; std::string err;
    IoTrace tr;
    if (input_mode == 0) {
        if (!read_alignments(argv[optind], h, rec, err)) fatal("update_gtf", err);
    } else {
        Records dummy;
        if (gtf_bam.empty()) fatal("update_gtf", "Couldn't read header of provided BAM file.\n");
        if (!read_alignments(gtf_bam, h, dummy, err)) fatal("update_gtf", err);
        if (!read_gtf(argv[optind], h, chains, true, err)) fatal("read_gtf_trans", err);
    }
    cn.seed(h);
    tr.lap("alignments");
    logf("read_anno_trans", (std::string("reading transcript annotation from ") + argv[optind + 1] + " ...\n").c_str());
    if (!read_gtf(argv[optind + 1], h, anno, false, err)) fatal("read_anno_trans", err);
    tr.lap("annotation gtf");
    logf("read_anno_trans", (std::string("reading transcript annotation from ") + argv[optind + 1] + " done.\n").c_str());
    SjTable sj;
    if (!sj_fn.empty() && !read_sj(sj_fn, cn, sj, err)) fatal("update_gtf", err);

    tr.lap("sj table");
    lrb_anno av = anno.view(); lrb_sj sv = sj.view();
    int rc = eng.set_tables(eng.self, &av, nullptr, sj.tid.empty() ? nullptr : &sv);
    if (rc) engine_fail(eng, "update_gtf", rc);
    tr.lap("tables upload");
    lrb_batch b = rec.view(); lrb_chains ch = chains.chains();
    RowNames rn; if (input_mode == 0) rn.rec = &rec; else rn.chains = &chains;
    const bool per_read_outputs = bam_gtf_fp || detail_fp || known_fp || novel_fp || unrecog_fp;
    if (!per_read_outputs && eng.update_table) {      // -o / -y / -E only: fetch just the rows that get printed
        lrb_trans_table tab{}; lrb_bed_list bed{}; int32_t counts[LRB_S_COUNT] = {0};
        rc = eng.update_table(eng.self, input_mode == 0 ? &b : nullptr, input_mode == 0 ? nullptr : &ch, &ep, &up, &tab, up.want_summary ? &bed : nullptr, counts);
        if (rc) engine_fail(eng, "update_gtf", rc);
        tr.lap("engine");
        emit_update_table(tab, up.want_summary ? &bed : nullptr, counts, rn, anno, h, cn, src.c_str(), anno.gene_n, (int)anno.n(), out_fp, summary_fp, bed_fp);
    } else {
        lrb_update_result res;
        rc = eng.update(eng.self, input_mode == 0 ? &b : nullptr, input_mode == 0 ? nullptr : &ch, &ep, &up, &res);
        if (rc) engine_fail(eng, "update_gtf", rc);
        tr.lap("engine");
        emit_update_outputs(res, rn, anno, h, cn, src.c_str(), anno.gene_n, (int)anno.n(),
                            out_fp, bam_gtf_fp, detail_fp, known_fp, novel_fp, unrecog_fp, summary_fp, bed_fp);
    }
    FILE *fps[] = {out_fp, bed_fp, bam_gtf_fp, detail_fp, known_fp, novel_fp, unrecog_fp, summary_fp};
    for (FILE *f : fps) if (f && f != stdout) fclose(f);
    tr.lap("emit");
    return 0;
}

// -------------------------------------------------------------------------------------------- unique-gtf
static int unique_usage()
{
    fprintf(stderr, "\n");
    fprintf(stderr, "Usage:   %s unique-gtf [option] <in.sorted.bam/in.sorted.gtf> > unique.gtf\n\n", PROG);
    fprintf(stderr, "Notice:  the BAM and GTF files should be sorted in advance.\n\n");
    fprintf(stderr, "Input options:\n\n");
    fprintf(stderr, "         -m --input-mode  [STR]    format of input file <in.bam/in.gtf>, BAM file(b) or GTF file(g). [b]\n");
    fprintf(stderr, "         -b --bam         [STR]    for GTF input <in.gtf>, BAM file is needed to obtain BAM header information. [NULL]\n");
    fprintf(stderr, "\n");
    fprintf(stderr, "Function options:\n\n");
    fprintf(stderr, "         -s --force-strand         force to match strand when merging transcripts. [False]\n");
    fprintf(stderr, "         -e --min-exon    [INT]    minimum length of internal exon. [%d]\n", 3);
    fprintf(stderr, "         -i --min-intron  [INT]    minimum length of intron. [%d]\n", 3);
    fprintf(stderr, "         -t --max-delet   [INT]    maximum length of deletion, longer deletion will be considered as intron. [%d]\n", 50);
    fprintf(stderr, "         -d --distance    [INT]    consider same if distance between two splice site is not bigger than d. [%d]\n", 0);
    fprintf(stderr, "         -D --DISTANCE    [INT]    consider same if distance between two start/end site is not bigger than D. [%d]\n", 0x7fffffff);
    fprintf(stderr, "         -f --frac        [INT]    consider same if overlapping between two single-exon transcript is bigger than f. [%.2f]\n", 0.80);
    fprintf(stderr, "\n");
    fprintf(stderr, "Output options:\n\n");
    fprintf(stderr, "         -I --intersect            output intersected transcript. [False]\n");
    fprintf(stderr, "         -o --output      [STR]    unique GTF file. [stdout]\n");
    fprintf(stderr, "         -S --source      [STR]    \'source\' field in GTF: program, database or project name. [%s]\n", PROG);
    fprintf(stderr, "\n");
    return 1;
}

static int cmd_unique(int argc, char **argv, Engine &eng)
{
    static const struct option lo[] = {
        {"input-mode", 1, NULL, 'm'}, {"bam", 1, NULL, 'b'}, {"force-strand", 0, NULL, 's'}, {"min-exon", 1, NULL, 'e'},
        {"min-intron", 1, NULL, 'i'}, {"distance", 1, NULL, 'd'}, {"DISTANCE", 1, NULL, 'D'}, {"frac", 1, NULL, 'f'},
        {"intersect", 0, NULL, 'I'}, {"output", 1, NULL, 'o'}, {"source", 1, NULL, 's'}, {0, 0, 0, 0}};
    lrb_update_params up; default_update_params(up);
    lrb_exon_params ep = {3, 3, 50};
    int input_mode = 0, intersect = 0; std::string gtf_bam, src = PROG; FILE *out_fp = stdout; int c;
    while ((c = getopt_long(argc, argv, "m:b:se:i:Id:D:f:o:S:", lo, NULL)) >= 0) {
        switch (c) {
        case 'm': if (optarg[0] == 'b') input_mode = 0; else if (optarg[0] == 'g') input_mode = 1; else return unique_usage(); break;
        case 'b': gtf_bam = optarg; { FILE *t = fopen(optarg, "rb"); if (!t) fatal("unique_gtf", std::string("Cannot open \"") + optarg + "\"\n"); fclose(t); } break;
        case 's': up.force_strand = 1; break;
        case 'e': ep.min_exon = atoi(optarg); break;
        case 'i': ep.min_intron = atoi(optarg); break;
        case 't': ep.max_delet = atoi(optarg); break;
        case 'd': up.ss_dis = atoi(optarg); break;
        case 'D': up.end_dis = atoi(optarg); break;
        case 'f': up.single_exon_ovlp_frac = (float)atof(optarg); break;
        case 'I': intersect = 1; break;
        case 'o': out_fp = fopen(optarg, "w"); break;
        case 'S': src = optarg; break;
        default: fprintf(stderr, "Error: unknown option: %s.\n", optarg); return unique_usage();
        }
    }
    if (argc - optind != 1) return unique_usage();
    Header h; Records rec; Anno chains; ChrNames cn; std::string err;
    if (input_mode == 0) {
        if (!read_alignments(argv[optind], h, rec, err)) fatal("unique_gtf", err);
    } else {
        Records dummy;
        if (gtf_bam.empty()) fatal("unique_gtf", "Couldn't read header of provided BAM file.\n");
        if (!read_alignments(gtf_bam, h, dummy, err)) fatal("unique_gtf", err);
        if (!read_gtf(argv[optind], h, chains, true, err)) fatal("read_gtf_trans", err);
    }
    cn.seed(h);
    lrb_unique_result res; lrb_batch b = rec.view(); lrb_chains ch = chains.chains();
    int rc = eng.unique(eng.self, input_mode == 0 ? &b : nullptr, input_mode == 0 ? nullptr : &ch, &ep, &up, &res);
    if (rc) engine_fail(eng, "unique_gtf", rc);
    RowNames rn; if (input_mode == 0) rn.rec = &rec; else rn.chains = &chains;
    emit_unique(out_fp, res, rn, cn, src.c_str(), intersect != 0);
    if (out_fp != stdout) fclose(out_fp);
    return 0;
}

int cli_main(int argc, char **argv, Engine &eng)
{
    if (argc < 2) return usage();
    if (strcmp(argv[1], "filter") == 0) return cmd_filter(argc - 1, argv + 1, eng);
    else if (strcmp(argv[1], "update-gtf") == 0) return cmd_update(argc - 1, argv + 1, eng);
    else if (strcmp(argv[1], "unique-gtf") == 0) return cmd_unique(argc - 1, argv + 1, eng);
    else if (strcmp(argv[1], "bam2gtf") == 0) return cmd_bam2gtf(argc - 1, argv + 1, eng);
    else if (strcmp(argv[1], "bam2sj") == 0) return cmd_bam2sj(argc - 1, argv + 1, eng);
    else if (strcmp(argv[1], "sort-gtf") == 0) return cmd_sort_gtf(argc - 1, argv + 1, eng);       // src/sort_gtf.sh
    else if (strcmp(argv[1], "fusion") == 0) {
        fprintf(stderr, "[main] command '%s' is outside the accelerated path of this build (see DESIGN.md); use the reference binary\n", argv[1]);
        return 1;
    }
    fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]);
    return 1;
}

}  // namespace lrb
