"""Motif enrichment between input and control regions (reference motifscan/stats.py:18-45).

Per motif the reference counts the regions that carry at least one site in each set, forms the
2 x 2 table [[n_in, N_in - n_in], [n_ctl, N_ctl - n_ctl]], runs the two one-sided Fisher exact
tests and Bonferroni-corrects the smaller p-value by the number of motifs.  Only `len()` of the
per-region site lists enters, so the counts come straight from the scan's arrays
(`MotifSites.regions_with_sites`, computed from the sorted site keys) and the tests are the same
`scipy.stats.fisher_exact` calls the reference makes: identical counts give identical p-values.
"""
from collections import namedtuple

import numpy as np
from scipy.stats import fisher_exact

EnrichmentResult = namedtuple(
    "EnrichmentResult",
    ["name", "n_input", "n_control", "fold_change", "p_enriched", "p_depleted", "p_corrected"])


def _hit_regions(motif_sites):
    """(regions with >= 1 site per motif, number of regions) for a MotifSites view or the
    reference's nested list [motif][region] -> sites."""
    if hasattr(motif_sites, "regions_with_sites"):
        return np.asarray(motif_sites.regions_with_sites(), dtype=np.int64), int(motif_sites.n_seqs)
    hits = np.array([sum(1 for cell in per_motif if len(cell) > 0) for per_motif in motif_sites],
                    dtype=np.int64)
    n_regions = len(motif_sites[0]) if len(motif_sites) else 0
    return hits, n_regions


def enrichment_from_counts(names, n_input, n_input_total, n_control, n_control_total):
    """The arithmetic of stats.py:32-43 on per-motif hit counts."""
    n_motifs = len(names)
    out = []
    for name, a, c in zip(names, n_input, n_control):
        a, c = int(a), int(c)
        if n_input_total > 0 and c > 0:
            fold = a * n_control_total / c / n_input_total            # stats.py:33
        else:
            fold = np.nan
        table = [[a, n_input_total - a], [c, n_control_total - c]]
        p_up = fisher_exact(table, "greater")[1]
        p_down = fisher_exact(table, "less")[1]
        out.append(EnrichmentResult(name, a, c, fold, p_up, p_down, min(min(p_up, p_down) * n_motifs, 1)))
    return out


def motif_enrichment(pwms, motif_sites, motif_sites_control):
    """Drop-in for `motifscan.stats.motif_enrichment`: list of EnrichmentResult in motif order."""
    n_in, total_in = _hit_regions(motif_sites)
    n_ctl, total_ctl = _hit_regions(motif_sites_control)
    names = [f"{pwm.matrix_id},{pwm.name}" for pwm in pwms]
    return enrichment_from_counts(names, n_in, total_in, n_ctl, total_ctl)
