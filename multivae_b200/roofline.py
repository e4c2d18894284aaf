"""Algorithmic work per launch of each native entry point (the numerators of bench.py's roofline); the
formulas are the ones stated in DESIGN.md.  Tensor-core kernels count USEFUL flops (real pixels only — the
halo rows the kernels also multiply are overhead, not work)."""

RIDGE_FLOP_PER_BYTE = 208.0   # measured sustained bf16 peak / measured HBM copy bandwidth (MEASURED_PEAKS.json)

# decoder layer tags (multivae_b200/nn/resnet_native.py) -> (H, Cin, Cout, taps)
_DEC_LAYERS = {
    "b1.sc": (7, 256, 128, 1), "b1.c0": (7, 256, 128, 9), "b1.c1": (7, 128, 128, 9),
    "b2.sc": (14, 128, 64, 1), "b2.c0": (14, 128, 64, 9), "b2.c1": (14, 64, 64, 9),
    "b3.c0": (28, 64, 64, 9), "b3.c1": (28, 64, 64, 9), "head": (28, 64, 3, 9),
}


# encoder layer tags: one branch of one modality, n_img = B images
_ENC_LAYERS = {
    "e.img": (28, 3, 64, 9), "e1.c0": (28, 64, 64, 9), "e1.c1": (28, 64, 64, 9),
    "e2.sc": (14, 64, 128, 1), "e2.c0": (14, 64, 64, 9), "e2.c1": (14, 64, 128, 9),
    "e3.sc": (7, 128, 256, 1), "e3.c0": (7, 128, 128, 9), "e3.c1": (7, 128, 256, 9),
}


def _layer(tag):
    base = tag[:-1] if tag.endswith("d") and tag[:-1] in _DEC_LAYERS else tag   # "b3.c1d" = data gradient of b3.c1
    if tag == "head.d":
        base = "head"
    return _DEC_LAYERS.get(base)


def describe(name, *, B, M, K, D, L, LW):
    """name = C-ABI symbol (optionally ':tag').  Returns {"bound", "work" (bytes or flops per launch)}."""
    sym, _, tag = name.partition(":")
    if "|" in tag:   # the wrapper stated this launch's work itself: "<layer>|f=<flops>|b=<bytes>"
        kv = dict(p.split("=") for p in tag.split("|")[1:])
        f, b = float(kv.get("f", 0)), float(kv.get("b", 0))
        # a contraction whose arithmetic intensity is below the ridge (~208 flop/B on this part) is an HBM kernel
        if f > 0 and b > 0 and f / b >= RIDGE_FLOP_PER_BYTE:
            return {"bound": "tensor", "work": f, "bytes": b}
        return {"bound": "hbm", "work": b, "flops": f}
    rows = M * K * B  # (cond modality, importance sample, batch sample) rows of one reconstructed modality
    if sym in ("mv_tapgemm", "mv_wgrad") and _layer(tag):
        H, cin, cout, taps = _layer(tag)
        return {"bound": "tensor", "work": 2.0 * rows * H * H * cin * cout * taps, "layer": (H, cin, cout, taps)}
    etag = tag[:-1] if tag.endswith("d") and tag[:-1] in _ENC_LAYERS else tag
    if sym in ("mv_tapgemm", "mv_wgrad", "mv_wgrad_slice") and etag in _ENC_LAYERS:
        H, cin, cout, taps = _ENC_LAYERS[etag]
        if sym == "mv_wgrad_slice":
            cout = 128
        return {"bound": "tensor", "work": 2.0 * B * H * H * cin * cout * taps, "layer": (H, cin, cout, taps)}
    if sym == "mv_moe_lpx_fwd":
        # read recon (bf16) once + targets (fp32) once, write lpx
        return {"bound": "hbm", "work": rows * D * 2 + B * D * 4 + rows * 4}
    if sym == "mv_moe_lpx_bwd":
        # read recon + targets + coef, write g_recon (bf16)
        return {"bound": "hbm", "work": rows * D * 2 * 2 + B * D * 4 + rows * 4}
    if sym == "mv_moe_lpx_fwd_multi":   # all M reconstructed modalities in one launch
        return {"bound": "hbm", "work": M * (rows * D * 2 + B * D * 4) + rows * 4}
    if sym == "mv_moe_lpx_bwd_multi":
        return {"bound": "hbm", "work": M * (rows * D * 2 * 2 + B * D * 4) + rows * 4}
    if sym == "mv_moe_lw_fwd":
        lat = M * K * B * (L + LW) * 4
        return {"bound": "hbm", "work": 2 * lat + 3 * rows * 4 + 4 * M * B * (L + LW) * 4 * 2}
    return {"bound": "hbm", "work": 0}
