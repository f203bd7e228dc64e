"""Golden VCF text produced by the REFERENCE'S OWN signature -> VCF code (build container only: needs /root/reference).

Imports focalsv/4_sv_calling/Dippav/extract_contig_signature_{CCS,CLR,ONT}.py unmodified and runs their top-level
extract_contig_sig_* on SAM files written by focalsv_b200.dropin.write_sam.  The imports the image lacks are stood in for:
matplotlib / tqdm by empty stubs, `pysam` by focalsv_b200.dropin (AlignmentFile + fetch: the two calls the reference makes),
`utils` is the reference's own utils.py.  Alignments of the "regions" cases come from the CPU oracle through
hook.realign_regions (test infrastructure), so that a GPU run of the same regions must reproduce the VCF byte for byte.
Output: tests/golden/vcf_golden.json.gz (records + VCF text per case; sequences are regenerated from seeds by tests/vcf_cases.py).
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
REFDIR = "/root/reference/focalsv/4_sv_calling/Dippav"

from focalsv_b200 import dropin, hook  # noqa: E402
import vcf_cases  # noqa: E402
from util import OracleRunner  # noqa: E402


def load_reference(platform):
    for name in ("matplotlib", "matplotlib.pyplot", "tqdm"):
        m = types.ModuleType(name)
        m.tqdm = lambda x, *a, **k: x
        sys.modules.setdefault(name, m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    shim = types.ModuleType("pysam")
    shim.AlignmentFile = dropin.AlignmentFile
    sys.modules["pysam"] = shim
    if REFDIR not in sys.path:
        sys.path.insert(0, REFDIR)            # `from utils import load_contigs` -> the reference's own utils.py
    sys.modules.pop("utils", None)
    spec = importlib.util.spec_from_file_location("ref_sig_" + platform, os.path.join(REFDIR, "extract_contig_signature_%s.py" % platform))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def write_fasta(path, items):
    with open(path, "w") as fh:
        for name, seq in items:
            fh.write(">%s\n" % name)
            for i in range(0, len(seq), 60):
                fh.write(seq[i:i + 60] + "\n")


def reference_vcf(platform, chrom_ascii, contigs_ascii, records):
    mod = load_reference(platform)
    with tempfile.TemporaryDirectory() as td:
        ref_fa, tig_fa, sam = os.path.join(td, "ref_chr21.fa"), os.path.join(td, "assemblies.fa"), os.path.join(td, "assemblies.sorted.sam")
        write_fasta(ref_fa, [(vcf_cases.CHROM, chrom_ascii)])
        write_fasta(tig_fa, sorted(contigs_ascii.items()))
        dropin.write_sam(sam, records, [(vcf_cases.CHROM, len(chrom_ascii))])
        fn = getattr(mod, "extract_contig_sig_" + platform)
        fn(chr_number=21, bam_path=sam, header_path=os.path.join(REFDIR, "header"), ref_path=ref_fa, contig_path=tig_fa, output_dir=td)
        return open(os.path.join(td, "dippav_variant_chr21.vcf")).read()


def main():
    from oracle import oracle as O
    out = {"source": REFDIR + "/extract_contig_signature_{CCS,CLR,ONT}.py (extract_contig_sig_*), header: " + REFDIR + "/header", "cases": []}
    header = open(os.path.join(REFDIR, "header")).read()
    out["header"] = header
    for kind, platform, seed in vcf_cases.CASES:
        if kind == "regions":
            chrom, windows, contigs = vcf_cases.region_case(seed, platform)
            aligned = hook.realign_regions(OracleRunner(O), windows, contigs, preset="asm5", bw=2000)
            recs = dropin.sam_records(aligned, windows, contigs)
            tigs = {n: vcf_cases.ascii_of(s) for n, s in contigs}
        else:
            chrom, tigs, recs = vcf_cases.record_case(seed)
        vcf = reference_vcf(platform, vcf_cases.ascii_of(chrom), tigs, recs)
        body = [ln for ln in vcf.split("\n") if ln and not ln.startswith("#")]
        out["cases"].append({"kind": kind, "platform": platform, "seed": seed,
                             "records": [[r.qname, r.flag, r.reference_name, r.pos, r.mapq, r.cigar] for r in recs], "vcf": vcf})
        print("%-8s %s seed %d: %d records -> %d VCF rows (%d DEL, %d INS, %d 1/1)" % (
            kind, platform, seed, len(recs), len(body), sum("SVTYPE=DEL" in b for b in body), sum("SVTYPE=INS" in b for b in body),
            sum(b.endswith("1/1") for b in body)))
    import gzip
    with gzip.open(os.path.join(HERE, "vcf_golden.json.gz"), "wt") as fh:
        json.dump(out, fh)
    print("wrote", os.path.join(HERE, "vcf_golden.json.gz"))


if __name__ == "__main__":
    main()
