"""motifscan_b200.motif -- mirrors the reference's `motifscan.motif` package layout for the
pieces on the scan path (the extension module `cscore` first of all)."""
