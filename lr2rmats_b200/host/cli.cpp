// cli.cpp -- drop-in `lr2rmats` command line above the C ABI.
//
// Same subcommands, option letters, long-option names (quirks included: `-M` takes an argument, `--source` maps to 's',
// bam2gtf's long names are exon-min / intron-len -- SURVEY.md App. A.5/A.9/D.1), defaults, output files and exit codes as
// main.c:37-49, bam_filter.c:98-164, bam2gtf.c:120-161, update_gtf.c:995-1117, unique_gtf.c:86-158.  The per-alignment
// work itself is done by the engine (the CUDA library in the product binary).
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <getopt.h>
#include <string>
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
    rc = eng.filter(eng.self, &b, &fp, &res);
    if (rc) engine_fail(eng, "bam_filter", rc);
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
    else if (strcmp(argv[1], "fusion") == 0 || strcmp(argv[1], "bam2sj") == 0) {
        fprintf(stderr, "[main] command '%s' is outside the accelerated path of this build (see DESIGN.md); use the reference binary\n", argv[1]);
        return 1;
    }
    fprintf(stderr, "[main] unrecognized command '%s'\n", argv[1]);
    return 1;
}

}  // namespace lrb
