#!/usr/bin/env python3
"""Command-line assembler: the reference's ``assemble.py`` interface over the B200 path.

Same flags (``-t/--time``, ``-m/--memory``, ``-c/--count_min_sketch``, ``-s/--stdout``,
``-k/--kmer_length``, ``-f/--filter_threshold`` -- so ``--filter`` keeps working as an
argparse abbreviation -- and ``-e/--error``; assemble.py:13-37), same stdin format (first line
the number of reads, then ``read`` or ``read1|read2|distance`` per line, pairing detected from
the first read line; :40-71), same stdout / ``./output/<time>_k_f_e.FASTQ`` formats (:74-99).
``-p/--paired`` is accepted as an assertion on the detected input kind (upstream has no such
flag; the north-star command line uses it).
"""
import argparse
import sys
import time
from pathlib import Path

from debruijn_graph import DeBruijnGraph, PairedDeBruijnGraph
from debug_graph import DebugDeBruijnGraph, DebugCMSDeBruijnGraph
from debug_graph import DebugPairedDeBruijnGraph, DebugCMSPairedDeBruijnGraph


class IOHandler:
    @staticmethod
    def read_args(argv=None):
        parser = argparse.ArgumentParser(
            description="Generates contigs (consensus regions) of a parent string given a set of "
                        "substrings or paired substrings. Takes in the number of reads or read-pairs in "
                        "the first line, then a read (\"read\") or read-pair (\"read1|read2|mean_dist\") "
                        "on each subsequent line. Outputs a file containing contigs to the subdirectory "
                        "./output by default.",
            formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        flag = parser.add_argument
        flag('-t', '--time', action='store_true',
             help="prints time at each stage during runtime to track program progression")
        flag('-m', '--memory', action='store_true', help="prints size of major data structures during runtime")
        flag('-c', '--count_min_sketch', action='store_true',
             help="uses a probabilistic data structure instead of a dictionary")
        flag('-s', '--stdout', action='store_true', help="switches output to stdout instead of file write")
        flag('-k', '--kmer_length', type=int, help="the k-value used to break down reads")
        flag('-f', '--filter_threshold', type=int, help="filter threshold for erroneous kmers")
        flag('-e', '--error', type=int,
             help="allowed error in paired distance for reconstruction with paired-reads")
        flag('-p', '--paired', action='store_true',
             help="assert that the input holds read-pairs (pairing is detected from the input)")
        return parser.parse_args(argv)

    @staticmethod
    def read_input(stream=None):
        """(reads, paired, distance, number of bases) from stdin; one bulk read instead of a
        ``readline`` per read, same parsing rules (assemble.py:45-71)."""
        stream = sys.stdin if stream is None else stream
        raw = IOHandler._raw_bytes(stream)
        if raw is not None:
            # plain ASCII input: parsed by libga_b200 into one symbol buffer + one length per read (pinned
            # host memory when a GPU is present) -- no Python string per read; ``reads`` decodes on demand
            import ga_ingest
            parsed = ga_ingest.parse(raw)
            if parsed is not None:
                return parsed
            text = raw.decode(getattr(stream, "encoding", None) or "utf-8", getattr(stream, "errors", None) or "strict")
            lines = text.replace("\r\n", "\n").replace("\r", "\n").split("\n")    # universal newlines, as text mode
        else:
            lines = stream.read().split("\n")
        wanted = int(lines[0].strip())
        body = lines[1:]

        def line(i):
            return body[i].strip() if i < len(body) else ""

        first = line(0).split('|')
        count = max(wanted, 1)              # upstream always consumes one read line
        if len(first) > 1:
            reads, distance, bases = [], 0, 0
            for i in range(count):
                read1, read2, distance = line(i).split('|')
                reads.append((read1, read2))
                bases += len(read1) + len(read2)
            return (reads, True, int(distance), bases)
        reads = [line(i) for i in range(count)]
        return (reads, False, 0, sum(map(len, reads)))

    @staticmethod
    def _raw_bytes(stream):
        """The whole input as bytes when the stream can give them (stdin's buffer, a binary file object)."""
        binary = getattr(stream, "buffer", None)
        if binary is not None:
            return binary.read()
        if isinstance(stream, (bytes, bytearray)):
            return bytes(stream)
        mode = getattr(stream, "mode", "")
        if isinstance(mode, str) and "b" in mode:
            return stream.read()
        import io
        if isinstance(stream, io.BytesIO):
            return stream.read()
        return None

    @staticmethod
    def _report(contigs, start_time, sep):
        yield ">Time started:" + sep + time.strftime("%c", time.localtime(start_time))
        yield ">Number of contigs:" + sep + str(len(contigs))
        for number, contig in enumerate(contigs, 1):
            yield ">CONTIG" + str(number)
            yield contig
        yield ">Time finished:" + sep + time.strftime("%c", time.localtime())

    @staticmethod
    def write_stdout(contigs, constants, start_time):
        # upstream prints with ``print(a, b)`` after a trailing space: two blanks after the colon
        sys.stdout.write("\n".join(IOHandler._report(contigs, start_time, "  ")) + "\n")

    @staticmethod
    def write_FASTQ(contigs, constants, start_time):
        out_dir = Path("./output")
        out_dir.mkdir(exist_ok=True)
        stamp = time.strftime("%b_%d_%H:%M:%S_%Y", time.localtime(start_time))
        name = "{0}_k{1}_f{2}_e{3}.FASTQ".format(stamp, *constants)
        with (out_dir / name).open(mode='x') as handle:      # no trailing newline, as upstream
            handle.write("\n".join(IOHandler._report(contigs, start_time, " ")))


class DebugIOHandler(IOHandler):
    @staticmethod
    def read_input(print_runtime=True, print_syssizeof=False, start_time=0):
        if print_runtime:
            print("\n>--- STARTING ASSEMBLY PROGRAM AT T = 0.00 ---")
        reads, paired, distance, bases = IOHandler.read_input()
        if print_runtime:
            print(">FINISHED READING INPUT AT T = {:.2f}".format(time.time() - start_time))
        if print_syssizeof:
            print(">SIZE OF READ CONTAINER: {:,}".format(sys.getsizeof(reads)))
            print(">SIZE OF ALL READ STRINGS: {:,}".format(sum(sys.getsizeof(r) for r in reads)))
        return (reads, paired, int(distance), bases)

    @staticmethod
    def write_FASTQ(contigs, constants, start_time, timeit):
        if timeit:
            print(">--- WRITING TO FILE AT T = {:.2f} ---".format(time.time() - start_time))
        IOHandler.write_FASTQ(contigs, constants, start_time)


def _constants(graph):
    return (graph.KMER_LEN, graph.HAMMING_DIST, graph.ALLOWED_PAIRED_DIST_ERROR)


def _expect_pairing(paired, expect_paired):
    if expect_paired and not paired:
        raise SystemExit("assemble.py: --paired given but the input holds unpaired reads")


def assemble_with_options(print_runtime=False, print_syssizeof=False, using_count_min_sketch=False,
                          start_time=None, k=None, hamming_dist=None, paired_error=None, expect_paired=False):
    start_time = time.time() if start_time is None else start_time
    reads, paired, _, _ = DebugIOHandler.read_input(
        print_runtime=print_runtime, start_time=start_time, print_syssizeof=print_syssizeof)
    _expect_pairing(paired, expect_paired)
    choice = {(False, False): DebugDeBruijnGraph, (False, True): DebugPairedDeBruijnGraph,
              (True, False): DebugCMSDeBruijnGraph, (True, True): DebugCMSPairedDeBruijnGraph}
    graph = choice[(bool(using_count_min_sketch), paired)](
        print_runtime=print_runtime, start_time=start_time, print_syssizeof=print_syssizeof,
        reads=reads, k=k, hamming_dist=hamming_dist, paired_error=paired_error)
    return graph.enumerate_contigs(), _constants(graph)


def assemble_without_options(k=None, hamming_dist=None, paired_error=None, expect_paired=False):
    reads, paired, _, _ = IOHandler.read_input()
    _expect_pairing(paired, expect_paired)
    cls = PairedDeBruijnGraph if paired else DeBruijnGraph
    graph = cls(reads, k=k, hamming_dist=hamming_dist, paired_error=paired_error)
    return graph.enumerate_contigs(), _constants(graph)


def main(argv=None):
    start_time = time.time()
    args = IOHandler.read_args(argv)
    if args.time or args.memory or args.count_min_sketch:
        contigs, constants = assemble_with_options(
            print_runtime=args.time, print_syssizeof=args.memory,
            using_count_min_sketch=args.count_min_sketch, start_time=start_time, k=args.kmer_length,
            hamming_dist=args.filter_threshold, paired_error=args.error, expect_paired=args.paired)
    else:
        contigs, constants = assemble_without_options(
            k=args.kmer_length, hamming_dist=args.filter_threshold, paired_error=args.error,
            expect_paired=args.paired)
    if args.stdout:
        IOHandler.write_stdout(contigs, constants, start_time)
    else:
        DebugIOHandler.write_FASTQ(contigs, constants, start_time, timeit=args.time)
    if args.time:
        print(">--- PROGRAM FINISHED AT T = {:.2f} ---".format(time.time() - start_time))


if __name__ == "__main__":
    main()
