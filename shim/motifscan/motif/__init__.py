"""Placeholder package so the shim resolves as motifscan.motif.cscore in tests (not part of the shim)."""
