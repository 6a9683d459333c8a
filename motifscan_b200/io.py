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


def _write_sites_table_blocks(f_num, f_score, regions, motif_sites, n_pwms, block_rows=8192):
    """The two tables block by block from the scan's arrays: the cells of a block of rows are filled and
    formatted by the library on host threads (msb_format_site_tables; numbers look like Python's str()), so
    200,000 regions x 1,900 motifs never exist as a dense table, let alone as Python objects."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    off = np.ascontiguousarray(motif_sites._off, dtype=np.int64)
    seq = np.ascontiguousarray(motif_sites._seq_idx, dtype=np.int32)
    score = np.ascontiguousarray(motif_sites._score, dtype=np.float64)
    if seq.size == 0:
        seq, score = np.zeros(1, np.int32), np.zeros(1, np.float64)
    n_threads = min(os.cpu_count() or 1, 32)
    for r0 in range(0, len(regions), block_rows):
        r1 = min(r0 + block_rows, len(regions))
        leads = [f"{g.chrom}\t{g.start + 1}\t{g.end}\t".encode() for g in regions[r0:r1]]
        lead_off = np.zeros(len(leads) + 1, dtype=np.int64)
        np.cumsum([len(x) for x in leads], out=lead_off[1:])
        p_num, p_score = ctypes.c_void_p(), ctypes.c_void_p()
        n_num, n_score = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(lib.msb_format_site_tables(n_pwms, _lib.ptr(off, ctypes.c_int64), _lib.ptr(seq, ctypes.c_int32),
                                              _lib.ptr(score, ctypes.c_double), r0, r1, b"".join(leads),
                                              _lib.ptr(lead_off, ctypes.c_int64), ctypes.byref(p_num), ctypes.byref(n_num),
                                              ctypes.byref(p_score), ctypes.byref(n_score), n_threads))
        try:
            # straight from the library's buffers to the files, no copy into a bytes object
            f_num.write((ctypes.c_char * n_num.value).from_address(p_num.value))
            f_score.write((ctypes.c_char * n_score.value).from_address(p_score.value))
        finally:
            lib.msb_text_free(p_num)
            lib.msb_text_free(p_score)


def write_sites_table(output_dir, pwms, regions, motif_sites):
    os.makedirs(output_dir, exist_ok=True)
    header = "chr\tstart\tend\t" + "\t".join(_motif_names(pwms)) + "\n"
    if hasattr(motif_sites, "_off"):     # a scan's arrays: streamed through the native formatter
        with open(os.path.join(output_dir, "motif_sites_number.xls"), "wb") as f_num, \
                open(os.path.join(output_dir, "motif_sites_score.xls"), "wb") as f_score:
            f_num.write(header.encode())
            f_score.write(header.encode())
            _write_sites_table_blocks(f_num, f_score, list(regions), motif_sites, len(pwms))
        return
    counts, best = _cell_tables(motif_sites, len(pwms), len(regions))
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
