"""motifscan.motif.cscore -- B200 backend (replaces the C extension built from motifscan/motif/cscore.c,
reference setup.py:53-56; same two callables, cscore.c:479-482)."""
from motifscan_b200.motif.cscore import c_score, c_scan_motif  # noqa: F401
