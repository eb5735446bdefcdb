"""Command line of pyrodigal_b200: the options of the reference CLI (src/pyrodigal/cli.py:64-323, itself an emulation
of the Prodigal command line), driving the GPU path.

Differences in mechanism, not in interface: the reference maps `find_genes` over the records of the input with a
thread or process pool (`-j`, `--pool`); here the whole file is parsed into one buffer (`fasta.read_batch`) and
every contig goes through ONE batched GPU call, so `-j` / `--pool` are accepted and ignored.  Outputs are written
by the same writers (`Genes.write_gff` ...), in input order."""
import argparse
import contextlib
import os
import sys
import warnings

from . import __version__
from . import fasta
from .lib import TRANSLATION_TABLES, GeneFinder, TrainingInfo


def argument_parser(prog="pyrodigal_b200", version=__version__, input_required=True,
                    formatter_class=argparse.ArgumentDefaultsHelpFormatter):
    p = argparse.ArgumentParser(prog=prog, add_help=False, formatter_class=formatter_class)
    p.add_argument("-a", metavar="trans_file", help="Write protein translations to the selected file.")
    p.add_argument("-c", action="store_true", default=False, help="Closed ends. Do not allow genes to run off edges.")
    p.add_argument("-d", metavar="nuc_file", help="Write nucleotide sequences of genes to the selected file.")
    p.add_argument("-f", metavar="output_type", choices={"gff", "gbk"}, default="gff", help="Select output format.")
    p.add_argument("-g", metavar="tr_table", type=int, choices=TRANSLATION_TABLES, default=11,
                   help="Specify a translation table to use.")
    p.add_argument("-i", metavar="input_file", required=input_required, help="Specify FASTA input file.")
    p.add_argument("-m", action="store_true", default=False,
                   help="Treat runs of N as masked sequence; don't build genes across them.")
    p.add_argument("-n", action="store_true", default=False,
                   help="Bypass Shine-Dalgarno trainer and force a full motif scan.")
    p.add_argument("-o", metavar="output_file", help="Specify output file.")
    p.add_argument("-p", metavar="mode", choices={"single", "meta"}, default="single", help="Select procedure.")
    p.add_argument("-s", metavar="start_file", help="Write all potential genes (with scores) to the selected file.")
    p.add_argument("-t", metavar="training_file",
                   help="Write a training file (if none exists); otherwise, read and use the specified training file.")
    p.add_argument("-j", "--jobs", type=int, default=1, metavar="jobs",
                   help="Accepted for compatibility: all sequences are processed in one batched GPU call.")
    p.add_argument("-h", "--help", action="help", help="Show this help message and exit.")
    p.add_argument("-V", "--version", action="version", version="{} v{}".format(prog, version),
                   help="Show version number and exit.")
    p.add_argument("--min-gene", type=int, default=90, help="The minimum gene length.")
    p.add_argument("--min-edge-gene", type=int, default=60, help="The minimum edge gene length.")
    p.add_argument("--max-overlap", type=int, default=60,
                   help="The maximum number of nucleotides that can overlap between two genes on the same strand. "
                        "This must be lower or equal to the minimum gene length.")
    p.add_argument("--no-stop-codon", action="store_true", default=False,
                   help="Disables translation of stop codons into star characters (*) for complete genes.")
    p.add_argument("--pool", choices=("thread", "process"), default="thread",
                   help="Accepted for compatibility (see -j).")
    p.add_argument("--device", type=int, default=0, help="CUDA device to run on.")
    return p


def main(argv=None, stdout=sys.stdout, stderr=sys.stderr, stdin=sys.stdin, *, gene_finder_factory=GeneFinder,
         argument_parser=argument_parser, formatter_class=argparse.ArgumentDefaultsHelpFormatter):
    parser = argument_parser(input_required=stdin.isatty(), formatter_class=formatter_class)
    args = parser.parse_args(argv)
    with contextlib.ExitStack() as ctx:
        try:
            nuc_file = None if args.d is None else ctx.enter_context(open(args.d, "w"))
            prot_file = None if args.a is None else ctx.enter_context(open(args.a, "w"))
            scores_file = None if args.s is None else ctx.enter_context(open(args.s, "w"))
            out_file = stdout if args.o is None else ctx.enter_context(open(args.o, "w"))

            training_info = None
            if args.t is not None:
                if args.p == "meta":
                    print("Error: cannot specify metagenomic sequence with a training file.", file=stderr)
                    return 1
                if os.path.exists(args.t):
                    with open(args.t, "rb") as f:
                        training_info = TrainingInfo.load(f)

            batch = fasta.read_batch(stdin.buffer if args.i is None and hasattr(stdin, "buffer") else (stdin if args.i is None else args.i),
                                     pinned=True)
            for seq_id in batch.ids:
                if not seq_id:
                    warnings.warn("Input file contains a sequence without identifier", stacklevel=2)

            gene_finder = gene_finder_factory(meta=args.p == "meta", closed=args.c, mask=args.m, training_info=training_info,
                                              min_gene=args.min_gene, min_edge_gene=args.min_edge_gene,
                                              max_overlap=args.max_overlap, device=args.device)
            if args.p == "single" and training_info is None:
                # every record of the input is a contig of one genome, like Prodigal (cli.py:267-279)
                training_info = gene_finder.train(*(batch.sequence(k) for k in range(len(batch))), force_nonsd=args.n,
                                                  translation_table=args.g)
                if args.t is not None and not os.path.exists(args.t):
                    with open(args.t, "wb") as f:
                        training_info.dump(f)

            preds = gene_finder.find_genes_batch(batch.flat, batch.offsets, want_nodes=scores_file is not None)
            for seq_id, genes in zip(batch.ids, preds):
                if args.f == "gff":
                    genes.write_gff(out_file, seq_id)
                else:
                    genes.write_genbank(out_file, seq_id)
                if nuc_file is not None:
                    genes.write_genes(nuc_file, seq_id)
                if prot_file is not None:
                    genes.write_translations(prot_file, seq_id, include_stop=not args.no_stop_codon)
                if scores_file is not None:
                    genes.write_scores(scores_file, seq_id)
        except Exception as err:
            print("Error: {}".format(err), file=stderr)
            raise
    return 0


if __name__ == "__main__":
    sys.exit(main())
