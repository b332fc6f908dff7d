"""Deterministic synthetic CPIU (counting-process information unit) tables.

Mirrors the schema the reference's data-prep utilities produce
(/root/reference/utilities/cpiu.R:567-577: one row per person-interval with ``int.n``,
``rt`` = interval length with a partial last interval, derived covariates), as specified in
SURVEY.md section 8(d):

* person follow-up F ~ U(0.5, 8) years, half-year intervals -> <= 16 rows per person;
  ``int.n`` = 1..k is the stratifier, ``rt`` = 0.5 except the last row (remainder, > 0);
* covariates: 1/3 baseline (constant within person), 1/3 time-varying (per-person random walk
  with carry-forward), 1/3 count / elapsed-time style (non-decreasing within person);
* event ~ Bernoulli(1 - exp(-lambda * rt)), lambda = exp(b0 + b'x) tuned to ~3-6 % events; the
  person's rows end at the first event; response coded 1.0 / 2.0.

Two cardinality modes: ``levels=L`` quantises every numeric column to at most L distinct values
(mode Q -- the only mode the O(m * #distinct) CPU reference can finish), ``levels=0`` keeps them
continuous (mode C).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

DATA_SEED = 20260101


@dataclass
class Cpiu:
    x: np.ndarray            # [p, n] float64: p contiguous columns of n (the .Call layout)
    y: np.ndarray            # [n] float64 in {1.0, 2.0}
    rt: np.ndarray           # [n] float64 risk time > 0
    strat: np.ndarray        # [n] int32 interval number >= 1
    person: np.ndarray       # [n] int32 person id (not an input of the path; for diagnostics)

    @property
    def n(self) -> int:
        return int(self.y.size)

    @property
    def p(self) -> int:
        return int(self.x.shape[0])


def _one_batch(rng: np.random.Generator, persons: int, p: int, interval: float, max_k: int):
    F = rng.uniform(interval, interval * max_k, size=persons)
    k = np.ceil(F / interval).astype(np.int64)
    k = np.minimum(k, max_k)
    start = np.zeros(persons + 1, np.int64)
    np.cumsum(k, out=start[1:])
    rows = int(start[-1])
    pid = np.repeat(np.arange(persons, dtype=np.int64), k)
    intn = (np.arange(rows, dtype=np.int64) - start[pid]) + 1
    rt = np.full(rows, interval)
    last = start[1:] - 1
    rem = F - interval * (k - 1)
    rt[last] = np.where(rem > 0, rem, interval)

    nb = (p + 2) // 3                 # baseline
    nt = (p + 1) // 3                 # time-varying
    nc = p - nb - nt                  # counts / elapsed
    x = np.empty((p, rows), np.float64)
    # baseline: age (continuous), sex / comorbidities (binary), a few ordinal scores
    for j in range(nb):
        kind = j % 4
        if kind == 0:
            v = rng.normal(60.0, 12.0, persons)
        elif kind == 1:
            v = (rng.random(persons) < 0.5).astype(np.float64)
        elif kind == 2:
            v = (rng.random(persons) < 0.2).astype(np.float64)
        else:
            v = rng.integers(0, 6, persons).astype(np.float64)
        x[j] = v[pid]
    # time-varying: random walk updated every `persist` intervals, carried forward in between
    for j in range(nt):
        persist = 1 + (j % 3)
        base = rng.normal(0.0, 1.0, persons)
        step = rng.normal(0.0, 0.35, rows)
        step[(intn - 1) % persist != 0] = 0.0
        step[intn == 1] = 0.0
        cs = np.cumsum(step)
        walk = cs - cs[start[pid]]
        x[nb + j] = base[pid] + walk
    # counts / elapsed time: non-decreasing within person
    for j in range(nc):
        if j % 2 == 0:
            inc = rng.poisson(0.3, rows).astype(np.float64)
            inc[intn == 1] = 0.0
            cs = np.cumsum(inc)
            x[nb + nt + j] = cs - cs[start[pid]]
        else:
            x[nb + nt + j] = (intn - 1) * interval + rng.integers(0, 3, persons)[pid] * interval
    return x, pid, intn, rt, start


def make_cpiu(n: int, p: int, *, levels: int = 64, seed: int = DATA_SEED, interval: float = 0.5,
              max_k: int = 16, event_rate: float = 0.05) -> Cpiu:
    """Exactly ``n`` rows x ``p`` covariates (the last person may be cut short)."""
    rng = np.random.default_rng(seed)
    beta = rng.normal(0.0, 0.35, p) * (rng.random(p) < 0.5)
    xs, ys, rts, sts, pids = [], [], [], [], []
    have = 0
    pid_off = 0
    while have < n:
        persons = max(64, int((n - have) / (0.5 * max_k) * 1.6) + 16)
        x, pid, intn, rt, start = _one_batch(rng, persons, p, interval, max_k)
        # standardise the linear predictor so the event rate is stable across p
        xc = x - x.mean(axis=1, keepdims=True)
        sd = x.std(axis=1, keepdims=True)
        sd[sd == 0] = 1.0
        eta = (beta[:, None] * (xc / sd)).sum(axis=0) / max(1.0, np.sqrt(np.count_nonzero(beta)))
        lam = np.exp(np.log(event_rate / interval) + eta - 0.5 * eta.var())
        ev = rng.random(rt.size) < (1.0 - np.exp(-lam * rt))
        # keep rows up to and including the person's first event
        cs = np.cumsum(ev)
        before = cs - ev - (cs - ev)[start[pid]]
        keep = before == 0
        xs.append(x[:, keep]); ys.append(ev[keep]); rts.append(rt[keep])
        sts.append(intn[keep]); pids.append(pid[keep] + pid_off)
        pid_off += persons
        have += int(keep.sum())
    x = np.concatenate(xs, axis=1)[:, :n]
    ev = np.concatenate(ys)[:n]
    rt = np.concatenate(rts)[:n]
    st = np.concatenate(sts)[:n]
    pid = np.concatenate(pids)[:n]
    if levels and levels > 0:
        x = quantise(x, levels)
    return Cpiu(x=np.ascontiguousarray(x), y=1.0 + ev.astype(np.float64), rt=np.ascontiguousarray(rt),
                strat=st.astype(np.int32), person=pid.astype(np.int32))


def quantise(x: np.ndarray, levels: int) -> np.ndarray:
    """Mode Q: map every column with more than ``levels`` distinct values onto ``levels``
    equal-width bins (value = bin midpoint); low-cardinality columns are left untouched."""
    out = np.empty_like(x)
    for j in range(x.shape[0]):
        col = x[j]
        lo, hi = col.min(), col.max()
        if hi <= lo or np.unique(col[: min(col.size, 200000)]).size <= levels and np.unique(col).size <= levels:
            out[j] = col
            continue
        w = (hi - lo) / levels
        b = np.minimum(np.floor((col - lo) / w), levels - 1)
        out[j] = lo + (b + 0.5) * w
    return out


def make_cpiu_torch(n: int, p: int, *, levels: int = 64, seed: int = DATA_SEED, interval: float = 0.5, max_k: int = 16,
                    event_rate: float = 0.05, device: str = "cuda", pin: bool = False) -> Cpiu:
    """The same schema as make_cpiu generated with torch on ``device`` (a `torch.Generator` stream, so the VALUES
    differ from the numpy generator's): for tables too large to synthesise on the host within a benchmark's
    budget (10M x 100 takes minutes in numpy, about a second on the GPU).  Returns host arrays."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    f64 = torch.float64
    U = lambda *s: torch.rand(*s, generator=g, device=device, dtype=f64)
    N = lambda *s: torch.randn(*s, generator=g, device=device, dtype=f64)
    beta = (N(p) * 0.35) * (U(p) < 0.5)
    persons = int(n / (0.5 * max_k) * 1.6) + 64
    while True:
        F = interval + U(persons) * (interval * max_k - interval)
        k = torch.clamp(torch.ceil(F / interval).to(torch.int64), max=max_k)
        start = torch.zeros(persons + 1, dtype=torch.int64, device=device)
        start[1:] = torch.cumsum(k, 0)
        rows = int(start[-1])
        pid = torch.repeat_interleave(torch.arange(persons, device=device), k)
        first = start[pid]
        intn = torch.arange(rows, device=device) - first + 1
        rt = torch.full((rows,), interval, dtype=f64, device=device)
        rem = F - interval * (k - 1).to(f64)
        rt[start[1:] - 1] = torch.where(rem > 0, rem, torch.full_like(rem, interval))
        nb, nt = (p + 2) // 3, (p + 1) // 3
        x = torch.empty((p, rows), dtype=f64, device=device)
        eta = torch.zeros(rows, dtype=f64, device=device)
        for j in range(p):
            if j < nb:                                             # baseline: constant within person
                kind = j % 4
                if kind == 0:
                    v = 60.0 + 12.0 * N(persons)
                elif kind == 1:
                    v = (U(persons) < 0.5).to(f64)
                elif kind == 2:
                    v = (U(persons) < 0.2).to(f64)
                else:
                    v = torch.floor(U(persons) * 6.0)
                col = v[pid]
            elif j < nb + nt:                                      # time-varying: random walk with carry-forward
                persist = 1 + ((j - nb) % 3)
                step = 0.35 * N(rows)
                step[((intn - 1) % persist != 0) | (intn == 1)] = 0.0
                cs = torch.cumsum(step, 0)
                col = N(persons)[pid] + (cs - cs[first])
            else:                                                  # counts / elapsed time: non-decreasing within person
                jj = j - nb - nt
                if jj % 2 == 0:
                    inc = torch.poisson(torch.full((rows,), 0.3, dtype=f64, device=device), generator=g)
                    inc[intn == 1] = 0.0
                    cs = torch.cumsum(inc, 0)
                    col = cs - cs[first]
                else:
                    col = (intn - 1).to(f64) * interval + torch.floor(U(persons) * 3.0)[pid] * interval
            x[j] = col
            if float(beta[j]) != 0.0:
                sd = col.std(unbiased=False)
                eta += beta[j] * (col - col.mean()) / (sd if float(sd) > 0 else 1.0)
        eta /= max(1.0, float(torch.count_nonzero(beta)) ** 0.5)
        lam = torch.exp(torch.log(torch.tensor(event_rate / interval, dtype=f64, device=device)) + eta - 0.5 * eta.var(unbiased=False))
        ev = U(rows) < (1.0 - torch.exp(-lam * rt))
        cs = torch.cumsum(ev.to(torch.int64), 0)
        before = cs - ev.to(torch.int64)
        keep = (before - before[first]) == 0                        # rows up to and including the person's first event
        idx = torch.nonzero(keep).squeeze(1)
        if idx.numel() >= n:
            idx = idx[:n]
            break
        persons = int(persons * 1.3)
    if levels and levels > 0:
        for j in range(p):
            col = x[j, idx]
            lo, hi = float(col.min()), float(col.max())
            if hi > lo and torch.unique(col).numel() > levels:
                w = (hi - lo) / levels
                b = torch.clamp(torch.floor((col - lo) / w), max=levels - 1)
                x[j, idx] = lo + (b + 0.5) * w
    xh = torch.empty((p, n), dtype=f64, pin_memory=(pin and device != "cpu"))
    for j in range(p):
        xh[j].copy_(x[j, idx])
    out = Cpiu(x=xh.numpy(), y=(1.0 + ev[idx].to(f64)).cpu().numpy(), rt=rt[idx].cpu().numpy(),
               strat=intn[idx].to(torch.int32).cpu().numpy(), person=pid[idx].to(torch.int32).cpu().numpy())
    return out
