"""Algorithmic work per launch of each native entry point (the numerators of bench.py's roofline);
the formulas are the ones stated in DESIGN.md."""


def describe(name, *, B, M, K, D, L, LW):
    """name = C-ABI symbol (optionally ':tag').  Returns {"bound", "work" (bytes or flops per launch)}."""
    sym = name.split(":")[0]
    rows = M * K * B  # (cond modality, importance sample, batch sample) rows of one reconstructed modality
    if sym == "mv_moe_lpx_fwd":
        # read recon (bf16) once + targets (fp32) once, write lpx
        return {"bound": "hbm", "work": rows * D * 2 + B * D * 4 + rows * 4}
    if sym == "mv_moe_lpx_bwd":
        # read recon + targets + coef, write g_recon (bf16)
        return {"bound": "hbm", "work": rows * D * 2 * 2 + B * D * 4 + rows * 4}
    if sym == "mv_moe_lw_fwd":
        lat = M * K * B * (L + LW) * 4
        return {"bound": "hbm", "work": 2 * lat + 3 * rows * 4 + 4 * M * B * (L + LW) * 4 * 2}
    return {"bound": "hbm", "work": 0}
