"""wfa-go compatible command line over AlignBatch (reference wfa-go/wfa-go.go:36-190).

Same flags, input format and report text as the reference CLI; all pairs of the
input file go to the GPU in one batched call instead of one Align per pair.

    python -m wfa_b200.cli [options] <query seq> <target seq>
    python -m wfa_b200.cli [options] -i input.txt
"""
import argparse
import sys

VERSION = "0.4.0"

USAGE = """
WFA alignment in Golang

 Author: Wei Shen <shenwei356@gmail.com>
   Code: https://github.com/shenwei356/wfa
Version: v%s (libwfacuda backend)

Input file format:
  see https://github.com/smarco/WFA-paper?tab=readme-ov-file#41-introduction-to-benchmarking-wfa-simple-tests
  Example:
  >ATTGGAAAATAGGATTGGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTTCGTCGTCCTTACGTTTCCGGAAGGGAGTGGTTAGCTCGAAGCCCA
  <GATTGGAAAATAGGATGGGGTTTGTTTATATTTGGGTTGAGGGATGTCCCACCTTGTCGTCCTTACGTTTCCGGAAGGGAGTGGTTGCTCGAAGCCCA

Usage:
  1. Align two sequences from the positional arguments.

        %s [options] <query seq> <target seq>

  2. Align sequence pairs from the input file (described above).

        %s [options] -i input.txt

Options/Flags:
  -N    do not output alignment (for benchmark)
  -a    do not use adaptive reduction
  -g    do not use global alignment
  -h    print help message
  -i string
        input file.
  -t    only show the aligned region
"""


def format_report(result, q, t, trim=False):
    """The text block of wfa-go.go:121-137 for one AlignmentResult-like object
    (anything with CIGAR(), AlignmentText() and the result fields)."""
    Q, A, T = result.AlignmentText(q, t, trim)
    pct = float(result.Matches) / float(result.AlignLen) * 100 if result.AlignLen else float("nan")
    return ("query   %s\n        %s\ntarget  %s\ncigar   %s\n\n"
            "align-score : %d\n"
            "match-region: q[%d, %d]/%d vs t[%d, %d]/%d\n"
            "align-length: %d, matches: %d (%.2f%%), gaps: %d, gap regions: %d\n\n") % (
        Q.decode("latin-1"), A.decode("latin-1"), T.decode("latin-1"), result.CIGAR(trim),
        result.Score, result.QBegin, result.QEnd, len(q), result.TBegin, result.TEnd, len(t),
        result.AlignLen, result.Matches, pct, result.Gaps, result.GapRegions)


def main(argv=None):
    ap = argparse.ArgumentParser(add_help=False)
    ap.add_argument("-h", action="store_true")
    ap.add_argument("-i", default="")
    ap.add_argument("-g", action="store_true")
    ap.add_argument("-a", action="store_true")
    ap.add_argument("-N", action="store_true")
    ap.add_argument("-t", action="store_true")
    ap.add_argument("seqs", nargs="*")
    args = ap.parse_args(argv)
    app = "wfa-go"
    if args.h:
        sys.stderr.write(USAGE % (VERSION, app, app))
        return 0
    from . import api, datagen
    if args.i:
        try:
            pairs = datagen.read_pair_file(args.i)
        except OSError:
            sys.stderr.write("failed to read file: %s\n" % args.i)
            return 1
    else:
        if len(args.seqs) != 2:
            sys.stderr.write('if flag -i not given, please give me two sequences. type "%s -h" for help.\n' % app)
            return 1
        pairs = [(args.seqs[0].encode(), args.seqs[1].encode())]
    algn = api.New(api.DefaultPenalties, api.Options(not args.g))            # wfa-go.go:96-98
    if not args.a:
        algn.AdaptiveReduction(api.AdaptiveReductionOption(10, 50, 1))       # :100-106
    try:
        # one batched call; with output wanted the CIGAR strings and the three text lines are
        # rendered on the GPU as well (wfacuda_batch_render)
        if args.N:
            results, errs = algn.AlignBatch([p[0] for p in pairs], [p[1] for p in pairs])
        else:
            results, errs = algn.AlignBatchRendered([p[0] for p in pairs], [p[1] for p in pairs], args.t)
        for (q, t), r, e in zip(pairs, results, errs):
            if e is not None:                                                # checkError, :185-190
                sys.stderr.write("%s\n" % e)
                return 1
            if not args.N:
                sys.stdout.write(format_report(r, q, t, args.t))
    finally:
        api.RecycleAligner(algn)
    return 0


if __name__ == "__main__":
    sys.exit(main())
