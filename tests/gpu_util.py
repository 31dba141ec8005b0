"""Helpers for the -m gpu tests: torch is used only for device memory and streams."""
import numpy as np
import torch

import coregex_b200 as cg


def dev_corpus(kind, seed, nbytes, first_block=0, literals=None):
    """Generate a synthetic corpus directly in HBM; returns a uint8 cuda tensor."""
    bs = cg.SYNTH_BLOCK[kind]
    assert nbytes % bs == 0
    t = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")[:nbytes]
    lit_ptr = off_ptr = 0
    keep = None
    nlit = 0
    if literals:
        blob, offs, nlit = cg._pack_literals(literals)
        dl = torch.from_numpy(blob).cuda()
        do = torch.from_numpy(offs).cuda()
        lit_ptr, off_ptr, keep = dl.data_ptr(), do.data_ptr(), (dl, do)
    cg.synth_device(kind, seed, t.data_ptr(), nbytes, first_block, lit_ptr, off_ptr, nlit)
    torch.cuda.synchronize()
    return t


def scan_device(regex, t, mode=cg.MODE_FINDALL, cap=None, base=0):
    """Device-resident scan of tensor t. Returns (total, flag, pairs ndarray or None)."""
    n = t.numel()
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    out = None
    out_ptr = 0
    if mode == cg.MODE_FINDALL:
        if cap is None:
            cap = max(1024, n // 8)
        out = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
        out_ptr = out.data_ptr()
    else:
        cap = 0
    regex.scan_device(t.data_ptr(), n, mode, out_ptr, cap, res.data_ptr(), base)
    torch.cuda.synchronize()
    total, flag = int(res[0].item()), int(res[1].item())
    pairs = None
    if out is not None:
        pairs = out[: min(total, cap)].cpu().numpy()
    return total, flag, pairs
