"""`motifscan scan` and `motifscan motif --build` over the CUDA scan path.

Same flags as the reference's subcommands (cli/main.py:407-430, 504-579) and the same outputs
(cli/scan.py:24-108, cli/motif.py:101-155).  Differences, all outside the scan path:
  * `-g` / `-m` are resolved through `~/.motifscanrc` ([genome] / [motif] sections, the
    reference's registry format, config.py:15-117) or taken as a directory path when one exists;
  * `--loc`, `--plot` and the install / list / search subcommands (gene annotation, matplotlib,
    remote databases) are out of scope and exit with a message;
  * `-t` is accepted and ignored: the device has no thread count; `--gpus N` scans on the first N
    GPUs of the box (one host thread and one context per GPU, results gathered in region order);
  * control regions are always whole-background random (`generate_control_regions` without genes):
    the gene-distance-matched branch of cli/scan.py needs the annotation layer, which is out of
    scope -- when the genome directory holds an annotation file a warning says so;

    python -m motifscan_b200 scan -i peaks.bed -m <motif set> -g <genome> -o out_dir
    python -m motifscan_b200 motif --build <motif set> -g <genome>
"""
import argparse
import logging
import os
import sys
from configparser import ConfigParser

from . import io as msio
from .builder import build_cutoffs
from .genome import Genome
from .motif import MotifPwms, PositionWeightMatrix, pfm_to_pwm, read_jaspar_pfms
from .region import REGION_FORMATS, generate_control_regions, load_motifscan_regions
from .scanner import Scanner
from .stats import motif_enrichment

logger = logging.getLogger("motifscan")
RC_PATH = os.path.expanduser("~/.motifscanrc")


def _registry(section, name, rc_path=None):
    """Directory registered for `name` in the reference's rc file, or `name` itself if it is one."""
    if os.path.isdir(name):
        return os.path.abspath(name), os.path.basename(os.path.normpath(name))
    cfg = ConfigParser()
    cfg.read(rc_path or os.environ.get("MOTIFSCANRC", RC_PATH))
    if cfg.has_option(section, name):
        return cfg.get(section, name), name
    raise SystemExit(f"motifscan: error: {section} {name!r} not found (not a directory, not in the rc file)")


def load_genome(name):
    path, short = _registry("genome", name)
    return Genome(short, path=path)


PACKED_MIN_BP = 1 << 26    # scans of at least this many window bases go through the packed, device-resident genome
BUILD_RESIDENT_MIN_SAMPLES = 200000    # `motif --build` with at least this many background samples samples from it too


def packed_genome(genome, window_bp, device=0):
    """The genome's packed planes (`genome.PackedGenome`, 0.375 B/bp) for a scan of `window_bp` bases, or None to
    fetch region strings from the FASTA as the reference does (scanner.py:76-87).  A cache written next to the
    FASTA (`<name>.packed.*`, like the `.fai`) is always used; without one the genome is encoded on the GPU and
    the cache written when the scan is large enough to pay for it (one pass over the FASTA)."""
    from . import engine
    from .genome import PackedGenome
    prefix = os.path.join(genome.path, f"{genome.name}.packed") if genome.path else None
    if prefix and os.path.isfile(prefix + ".chroms.tsv"):
        try:
            pg = PackedGenome.load(prefix, genome=genome)
            if pg.chrom_sizes == genome.chrom_sizes:
                return pg
            logger.warning("packed genome cache does not match the FASTA index: ignored")
        except (OSError, ValueError) as err:
            logger.warning(f"packed genome cache unreadable ({err}): ignored")
    if window_bp < PACKED_MIN_BP:
        return None
    logger.info("Packing the genome (2 bits per base + N mask) on the device")
    pg = PackedGenome.from_genome(genome, engine.default_context(device))
    if prefix:
        try:
            pg.save(prefix)
        except OSError:
            pass
    return pg


def pwms_path(motif_dir, name, genome_name):
    return os.path.join(motif_dir, f"{name}_{genome_name}_pwms.motifscan")   # motif/__init__.py:22


def load_built_pwms(name, genome_name):
    motif_dir, short = _registry("motif", name)
    path = pwms_path(motif_dir, short, genome_name)
    if not os.path.isfile(path):
        raise SystemExit(f"motifscan: error: PWMs of motif set {short!r} are not built for genome "
                         f"{genome_name!r} (run: motif --build {name} -g {genome_name})")
    pwms = MotifPwms(name=short, genome=genome_name)
    pwms.read_motifscan_pwms(path)
    return pwms


def run_scan(args):
    if args.location is not None:
        raise SystemExit("motifscan scan: --loc needs the gene annotation layer, which is out of scope here")
    if args.plot_dist:
        raise SystemExit("motifscan scan: --plot needs matplotlib, which is out of scope here")
    logger.info("===== Loading data =====")
    genome = load_genome(args.genome)
    pwms = load_built_pwms(args.motif, genome.name)
    regions = load_motifscan_regions(args.input_file, args.input_format)
    logger.info("===== Scanning motifs =====")
    devices = _devices(args.n_gpus)
    window_bp = sum((args.window_size if args.window_size > 0 else r.end - r.start) for r in regions)
    source = packed_genome(genome, window_bp * (1 if args.no_enrich else 2), devices[0]) or genome
    common = dict(genome=source, window_size=args.window_size, strand=args.strand, p_value=args.p_value,
                  remove_dup=True, n_threads=args.n_threads, devices=devices)
    sites = Scanner(regions=regions, **common).scan_motifs(pwms)
    logger.info("Saving the result tables")
    msio.write_sites_table(args.output_dir, pwms, regions, sites)
    if args.report_site:
        msio.write_sites_bed(args.output_dir, pwms, regions, sites)
    if not args.no_enrich:
        logger.info("===== Motif Enrichment =====")
        if args.control_file:
            controls = load_motifscan_regions(args.control_file, args.control_format)
        else:
            if genome.path and any(f.endswith(("_gene_annotation.txt", "refGene.txt")) for f in os.listdir(genome.path)):
                logger.warning("gene annotation found but unused: control regions are drawn from the whole "
                               "background, not matched by gene distance as the reference does (out of scope)")
            controls = generate_control_regions(args.n_random, regions, genome.chrom_sizes, random_seed=args.seed)
        control_sites = Scanner(regions=controls, **common).scan_motifs(pwms)
        msio.write_enrich_table(args.output_dir, motif_enrichment(pwms, sites, control_sites))
    logger.info("===== MotifScan Finished =====")
    return 0


def run_motif(args):
    if not args.build:
        raise SystemExit("motifscan motif: only --build is provided here (install / list need the remote databases)")
    if not args.genome:
        raise SystemExit("motifscan motif --build: error: argument -g/--genome is required")
    genome = load_genome(args.genome)
    motif_dir, short = _registry("motif", args.build)
    pfm_path = os.path.join(motif_dir, f"{short}_pfms.jaspar")               # motif/__init__.py:21
    if not os.path.isfile(pfm_path):
        raise SystemExit(f"motifscan motif --build: error: {pfm_path} not found")
    logger.info("Converting motif PFMs to PWMs")
    pwms = MotifPwms(name=short, genome=genome.name)
    for matrix_id, name, pfm in read_jaspar_pfms(pfm_path):
        pwms.append(PositionWeightMatrix(pfm_to_pwm(pfm, genome.bg_freq), name=name, matrix_id=matrix_id))
    logger.info("Sampling background sequences and scoring them on the device")
    # a large sample is drawn from the device-resident genome: the reference's RNG sequence stays on the host, the N
    # count of every attempt and the sampled windows come out of HBM instead of one FASTA fetch per attempt
    source = genome
    if args.n_random * args.n_repeat >= BUILD_RESIDENT_MIN_SAMPLES or os.path.isfile(
            os.path.join(genome.path or "", f"{genome.name}.packed.chroms.tsv")):
        pg = packed_genome(genome, PACKED_MIN_BP)
        if pg is not None:
            from .genome import DeviceGenome
            source = DeviceGenome(pg)
    build_cutoffs(pwms, source, n_random=args.n_random, n_repeat=args.n_repeat, max_n=args.max_n, seed=args.seed)
    pwms.write_motifscan_pwms(pwms_path(motif_dir, short, genome.name))
    logger.info("Successfully built!")
    return 0


def _devices(n_gpus):
    """`--gpus N`: the first N CUDA devices of this process (the reference's `-t` pool, cli/main.py:427-430,
    565-568, becomes one host thread per GPU); more than there are is an error, not a silent clamp."""
    from . import _lib
    have = _lib.device_count()
    if n_gpus > have:
        raise SystemExit(f"motifscan: error: --gpus {n_gpus} but only {have} CUDA device(s) are visible")
    return list(range(n_gpus))


def _non_negative(text):
    value = int(text)
    if value < 0:
        raise argparse.ArgumentTypeError(f"expect a non-negative integer, got {text!r}")
    return value


def _positive(text):
    value = int(text)
    if value <= 0:
        raise argparse.ArgumentTypeError(f"expect a positive integer, got {text!r}")
    return value


def make_parser():
    parser = argparse.ArgumentParser(prog="motifscan", description="MotifScan scan path on B200 (motifscan_b200)")
    sub = parser.add_subparsers(dest="command", required=True)

    scan = sub.add_parser("scan", help="Scan input regions to detect motif occurrences.")
    scan.add_argument("-i", dest="input_file", required=True, metavar="FILE")
    scan.add_argument("-f", dest="input_format", choices=sorted(REGION_FORMATS), default="bed")
    scan.add_argument("-m", "--motif", dest="motif", required=True, metavar="NAME")
    scan.add_argument("-g", "--genome", dest="genome", required=True, metavar="GENOME")
    scan.add_argument("-p", dest="p_value", default="1e-4", choices=["1e-2", "1e-3", "1e-4", "1e-5", "1e-6"])
    scan.add_argument("--loc", dest="location", choices=["promoter", "distal"], default=None)
    scan.add_argument("--upstream", dest="upstream", type=_positive, default=4000)
    scan.add_argument("--downstream", dest="downstream", type=_positive, default=2000)
    scan.add_argument("-w", "--window-size", dest="window_size", type=_non_negative, default=1000)
    scan.add_argument("--strand", dest="strand", choices=["both", "+", "-"], default="both")
    scan.add_argument("--no-enrich", dest="no_enrich", action="store_true")
    scan.add_argument("--n-random", dest="n_random", type=_non_negative, default=5)
    scan.add_argument("--seed", dest="seed", type=int, default=None)
    scan.add_argument("-c", dest="control_file", metavar="FILE")
    scan.add_argument("--cf", dest="control_format", choices=sorted(REGION_FORMATS), default="bed")
    scan.add_argument("-t", "--threads", dest="n_threads", type=int, default=1)
    scan.add_argument("--gpus", dest="n_gpus", type=_positive, default=1, metavar="N",
                      help="GPUs of this box to scan on (regions are dealt to them in contiguous blocks)")
    scan.add_argument("-o", "--output-dir", dest="output_dir", required=True, metavar="DIR")
    scan.add_argument("--site", dest="report_site", action="store_true")
    scan.add_argument("--plot", dest="plot_dist", action="store_true")
    scan.add_argument("--verbose", action="store_true")
    scan.set_defaults(func=run_scan)

    motif = sub.add_parser("motif", help="Motif set commands (only --build here).")
    motif.add_argument("--build", dest="build", metavar="NAME")
    motif.add_argument("-g", "--genome", dest="genome", metavar="GENOME")
    motif.add_argument("--n-random", dest="n_random", type=int, default=1000000)
    motif.add_argument("--n-repeat", dest="n_repeat", type=_positive, default=1)
    motif.add_argument("--max-n", dest="max_n", type=int, default=0)
    motif.add_argument("--seed", dest="seed", type=int, default=None)
    motif.add_argument("-t", "--threads", dest="n_threads", type=int, default=1)
    motif.add_argument("--verbose", action="store_true")
    motif.set_defaults(func=run_motif)
    return parser


def main(argv=None):
    args = make_parser().parse_args(argv)
    logging.basicConfig(stream=sys.stderr, level=logging.DEBUG if args.verbose else logging.INFO,
                        format="%(message)s")
    return args.func(args)


if __name__ == "__main__":
    sys.exit(main())
