"""Parity at the sizes BASELINE.json names (SURVEY.md §8d, App. B hazard 11: 64-bit offsets).

Buffers of several GiB are generated in HBM; the CUDA result is compared with the CPU oracle on
regenerated 256 KB windows (same generator, same seed), with windows straddling and beyond the
4 GiB mark inside the buffer, plus whole-output properties that do not need the oracle
(ordered, non-overlapping, count == COUNT-mode total, every match inside the buffer)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import coregex_b200 as cg
from bench import LIT64, parity_windows
from gpu_util import dev_corpus

pytestmark = pytest.mark.gpu
GIB = 1 << 30
IP = r"\d+\.\d+\.\d+\.\d+"
LIT16 = [b"error", b"warning", b"fatal", b"critical", b"timeout", b"refused", b"denied", b"panic",
         b"overflow", b"invalid", b"missing", b"corrupt", b"expired", b"blocked", b"aborted", b"unknown"]


def _scan_all(r, t, base=0):
    n = t.numel()
    res = torch.zeros(2, dtype=torch.int64, device="cuda")
    r.scan_device(t.data_ptr(), n, cg.MODE_COUNT, 0, 0, res.data_ptr(), base)
    torch.cuda.synchronize()
    total = int(res[0].item())
    out = torch.empty((total + 16, 2), dtype=torch.int64, device="cuda")
    r.scan_device(t.data_ptr(), n, cg.MODE_FINDALL, out.data_ptr(), total + 16, res.data_ptr(), base)
    torch.cuda.synchronize()
    assert int(res[0].item()) == total
    return out, total


def _properties(out, total, base, n):
    p = out[:total]
    assert int(p[0, 0]) >= base and int(p[-1, 1]) <= base + n
    assert bool((p[:, 1] > p[:, 0]).all())
    assert bool((p[1:, 0] >= p[:-1, 1]).all())   # ascending and non-overlapping


def test_one_regex_scans_haystacks_of_different_sizes():
    """The look-back words and group accumulators of a regex survive from launch to launch (epochs,
    self-cleaning): a regex that scans haystacks of different sizes — the pieces of a pipelined host
    call — must find them in the same place every time, and a host call over several pieces must
    equal the device scan of the whole."""
    from oracle_lib import Oracle
    for pat, kind, lits in [(IP, cg.SYNTH_LOG, None), ("|".join(x.decode() for x in LIT16), cg.SYNTH_TEXT, LIT16)]:
        r = cg.Compile(pat)
        o = Oracle(pat)
        for blocks in [5000, 700, 9000, 64, 9000, 3, 20000]:
            n = blocks * 4096
            t = dev_corpus(kind, 77, n, literals=lits)
            out, total = _scan_all(r, t)
            w = min(n, 64 * 4096)
            hay = cg.synth_host(kind, 77, w, first_block=0, literals=lits)
            want = o.find_all(np.frombuffer(hay, dtype=np.uint8))
            got = out[:total].cpu().numpy()
            k = int(np.searchsorted(got[:, 0], w - 64))
            kw = int(np.searchsorted(want[:, 0], w - 64))
            assert np.array_equal(got[:k], want[:kw]), (pat, blocks)
    # host entry: 700 MiB go through the pipeline in pieces of different sizes
    n = 700 << 20
    hay = cg.synth_host(cg.SYNTH_LOG, 5, n)
    r = cg.Compile(IP)
    got = r.find_all_index_array(hay)
    t = torch.from_numpy(np.frombuffer(hay, dtype=np.uint8).copy()).cuda()
    out, total = _scan_all(r, t)
    assert total == len(got) and np.array_equal(out[:total].cpu().numpy(), got)


def test_ns_ip_regex_beyond_4gib():
    """North-star pattern on a 6 GiB shard that sits at block 1<<22 of the logical corpus: matches
    whose offset INSIDE the buffer exceeds 4 GiB are compared with the oracle."""
    n = 6 * GIB
    first_block = 1 << 22
    t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")[:n]
    cg.synth_device(cg.SYNTH_LOG, 0xC0FFEE, t.data_ptr(), n, first_block=first_block)
    r = cg.Compile(IP)
    out, total = _scan_all(r, t, base=first_block * 4096)
    _properties(out, total, first_block * 4096, n)
    par = parity_windows(cg, out, total, n, first_block, 16)
    assert par["ok"] and par["above_4gib"] >= 8 and par["windows"] >= 16


def test_c3_slim_teddy_4gb():
    n = 4 * GIB
    t = dev_corpus(cg.SYNTH_TEXT, 0xC0FFEE + 3, n, literals=LIT16)
    pat = b"|".join(LIT16).decode()
    r = cg.Compile(pat)
    assert r.strategy == "UseTeddy"
    out, total = _scan_all(r, t)
    _properties(out, total, 0, n)
    par = parity_windows(cg, out, total, n, 0, 12, seed=0xC0FFEE + 3, kind=cg.SYNTH_TEXT, pattern=pat, literals=LIT16)
    assert par["ok"] and par["windows"] >= 12


def test_c5_fat_teddy_8gb_shard():
    n = 8 * GIB
    first_block = 3 * (n // 4096)        # shard 3 of 8
    t = dev_corpus(cg.SYNTH_TEXT, 0xC0FFEE + 5, n, first_block=first_block, literals=LIT64)
    pat = b"|".join(LIT64).decode()
    r = cg.Compile(pat)
    out, total = _scan_all(r, t, base=first_block * 4096)
    _properties(out, total, first_block * 4096, n)
    par = parity_windows(cg, out, total, n, first_block, 16, seed=0xC0FFEE + 5, kind=cg.SYNTH_TEXT, pattern=pat,
                         literals=LIT64)
    assert par["ok"] and par["above_4gib"] >= 8
    # the compact wire format rebuilds exactly these pairs (device codec; segment table across 4 GiB)
    from coregex_b200 import shard
    wire, bad = shard.pack_offsets(out[:total], first_block * 4096, n)
    back = shard.unpack_offsets(wire, total, first_block * 4096, n)
    torch.cuda.synchronize()
    assert int(bad.item()) == 0 and torch.equal(back, out[:total])
    assert wire.numel() == shard._wire_layout(total, n)[1] <= 6 * total + 64


def test_wire_codec_device_equals_host_twin():
    from coregex_b200 import shard
    base, shard_len = 7 << 33, (9 << 30) + 4096
    rng = np.random.Generator(np.random.PCG64(5))
    starts = np.sort(rng.integers(0, shard_len - 70000, 200000)).astype(np.int64)
    starts = np.unique(np.concatenate([starts, [0, (1 << 32) - 1, 1 << 32, (2 << 32) - 1, 2 << 32]]))
    lens = rng.integers(0, 65536, starts.size).astype(np.int64)
    pairs = torch.from_numpy(np.stack([base + starts, base + starts + lens], axis=1))
    w_host = shard.pack_offsets(pairs, base, shard_len)
    w_dev, bad = shard.pack_offsets(pairs.cuda(), base, shard_len)
    assert int(bad.item()) == 0 and torch.equal(w_dev.cpu(), w_host)
    assert torch.equal(shard.unpack_offsets(w_dev, starts.size, base, shard_len).cpu(), pairs)
    long = torch.tensor([[base, base + 65536]], dtype=torch.int64).cuda()
    _, bad = shard.pack_offsets(long, base, shard_len)
    assert int(bad.item()) == 1
