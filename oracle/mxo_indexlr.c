/*
 * mxo_indexlr -- command-line front end of the CPU ORACLE (test infrastructure only).
 * Accepts the argv the reference passes to btllib's indexlr (ntJoin:205,
 * bin/ntjoin_utils.py:197-198):  [--seq] [--long] [--pos] [--strand] -k K -w W -t T [-o OUT] FASTA
 * plus oracle-only switches  --canonical sum|min  --tie right|left.
 */
#include "mxo.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int num_arg(int argc, char** argv, int* i, const char* flag, long* v)
{
    size_t fl = strlen(flag);
    if (strncmp(argv[*i], flag, fl)) return 0;
    if (argv[*i][fl]) { *v = atol(argv[*i] + fl); return 1; }
    if (*i + 1 >= argc) { fprintf(stderr, "mxo_indexlr: %s needs a value\n", flag); exit(2); }
    *v = atol(argv[++*i]);
    return 1;
}

int main(int argc, char** argv)
{
    long k = 0, w = 0, t = 1;
    int pos = 0, strand = 0, seq = 0, canonical = MXO_CANON_SUM, tie = MXO_TIE_RIGHT;
    const char *in = NULL, *outp = "-";
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--pos")) pos = 1;
        else if (!strcmp(argv[i], "--strand")) strand = 1;
        else if (!strcmp(argv[i], "--seq")) seq = 1;
        else if (!strcmp(argv[i], "--long") || !strcmp(argv[i], "--bx")) ;
        else if (!strcmp(argv[i], "--canonical") && i + 1 < argc)
            canonical = !strcmp(argv[++i], "min") ? MXO_CANON_MIN : MXO_CANON_SUM;
        else if (!strcmp(argv[i], "--tie") && i + 1 < argc)
            tie = !strcmp(argv[++i], "left") ? MXO_TIE_LEFT : MXO_TIE_RIGHT;
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) outp = argv[++i];
        else if (num_arg(argc, argv, &i, "-k", &k)) ;
        else if (num_arg(argc, argv, &i, "-w", &w)) ;
        else if (num_arg(argc, argv, &i, "-t", &t)) ;
        else if (argv[i][0] == '-' && argv[i][1]) { fprintf(stderr, "mxo_indexlr: unknown option %s\n", argv[i]); return 2; }
        else in = argv[i];
    }
    if (!in || k <= 0 || w <= 0) { fprintf(stderr, "usage: mxo_indexlr -k K -w W [-t T] [--pos] [--strand] [--seq] [-o OUT] FASTA\n"); return 2; }
    mxo_fasta_t fa;
    int rc = mxo_read_fasta(in, &fa);
    if (rc) { fprintf(stderr, "mxo_indexlr: cannot read %s (%d)\n", in, rc); return 1; }
    mxo_mx_t* mx = NULL; size_t n = 0;
    rc = mxo_sketch_buffers(fa.seq, fa.offsets, fa.n_contigs, (unsigned)k, (unsigned)w, canonical, tie, (int)t, &mx, &n);
    if (rc) { fprintf(stderr, "mxo_indexlr: sketch failed (%d)\n", rc); return 1; }
    rc = mxo_write_tsv(outp, &fa, mx, n, (unsigned)k, pos, strand, seq);
    mxo_free(mx);
    mxo_free_fasta(&fa);
    return rc ? 1 : 0;
}
