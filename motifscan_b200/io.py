"""Result writers of `motifscan scan` (reference motifscan/io/__init__.py:12-71): same file names,
columns and number formatting (`str()` of ints and Python floats), fed from the scan's arrays.

    motif_sites_number.xls   chr, 1-based start, end, then the site count per motif
    motif_sites_score.xls    same header, the best score per motif or `NA`
    motif_sites/<id>_<name>_sites.bed   chrom, site start, start + L, '.', score, strand
    motif_enrichment.xls     sorted by (enriched p-value, -fold change)
"""
import os
import re

import numpy as np


def replace_special_char(name):
    """Motif name as a file name (reference io/utils.py:15-16): `-`, `:`, `.`, `/`, `*` -> `_`."""
    return re.sub(r"[-:./*]", "_", name)


def _motif_names(pwms, sep=","):
    return [f"{pwm.matrix_id}{sep}{pwm.name}" for pwm in pwms]


def _cell_tables(motif_sites, n_pwms, n_regions):
    """(counts int64[n_regions, n_pwms], best float64[n_regions, n_pwms] with NaN where empty)."""
    counts = np.zeros((n_regions, n_pwms), dtype=np.int64)
    best = np.full((n_regions, n_pwms), np.nan)
    if hasattr(motif_sites, "_off"):     # MotifSites: sites sorted by (motif, region)
        for m in range(n_pwms):
            a, b = int(motif_sites._off[m]), int(motif_sites._off[m + 1])
            if b > a:
                seq = motif_sites._seq_idx[a:b]
                np.add.at(counts[:, m], seq, 1)
                col = np.full(n_regions, -np.inf)
                np.maximum.at(col, seq, motif_sites._score[a:b])
                hit = counts[:, m] > 0
                best[hit, m] = col[hit]
    else:
        for m, per_motif in enumerate(motif_sites):
            for r, cell in enumerate(per_motif):
                counts[r, m] = len(cell)
                if cell:
                    best[r, m] = max(site.score for site in cell)
    return counts, best


def write_sites_table(output_dir, pwms, regions, motif_sites):
    os.makedirs(output_dir, exist_ok=True)
    counts, best = _cell_tables(motif_sites, len(pwms), len(regions))
    header = "chr\tstart\tend\t" + "\t".join(_motif_names(pwms)) + "\n"
    with open(os.path.join(output_dir, "motif_sites_number.xls"), "w") as f_num, \
            open(os.path.join(output_dir, "motif_sites_score.xls"), "w") as f_score:
        f_num.write(header)
        f_score.write(header)
        for r, region in enumerate(regions):
            lead = f"{region.chrom}\t{region.start + 1}\t{region.end}\t"
            f_num.write(lead + "\t".join(map(str, counts[r].tolist())) + "\n")
            row = best[r].tolist()
            f_score.write(lead + "\t".join("NA" if c == 0 else str(v) for c, v in zip(counts[r].tolist(), row)) + "\n")


def write_sites_bed(output_dir, pwms, regions, motif_sites):
    out_dir = os.path.join(output_dir, "motif_sites")
    os.makedirs(out_dir, exist_ok=True)
    for pwm, per_motif in zip(pwms, motif_sites):
        path = os.path.join(out_dir, replace_special_char(f"{pwm.matrix_id}_{pwm.name}") + "_sites.bed")
        with open(path, "w") as out:
            for region, cell in zip(regions, per_motif):
                for site in cell:
                    out.write(f"{region.chrom}\t{site.start}\t{site.start + pwm.length}\t.\t{site.score}\t{site.strand}\n")


def write_enrich_table(output_dir, enrichment_results):
    os.makedirs(output_dir, exist_ok=True)
    enrichment_results.sort(key=lambda res: (res.p_enriched, -res.fold_change))   # io/__init__.py:63
    with open(os.path.join(output_dir, "motif_enrichment.xls"), "w") as out:
        out.write("Motif\tNum_input_regions\tNum_control_regions\tFold_change\tEnriched_P_value\t"
                  "Depleted_P_value\tCorrected_P_value\n")
        for res in enrichment_results:
            out.write(f"{res.name}\t{res.n_input}\t{res.n_control}\t{res.fold_change}\t{res.p_enriched}\t"
                      f"{res.p_depleted}\t{res.p_corrected}\n")
