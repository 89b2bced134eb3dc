"""Deterministic synthetic trigenic datasets.

The reference ships no readable data (every data file is a git-LFS stub), so tests, the
golden-vector generator and bench.py all draw their inputs from here.  Labels come from a
planted mixed-membership block model so that AUC is informative.

Two triplet shapes:
  * ``uniform``  - every triple of distinct genes equally likely.
  * ``kuzmin``   - (query pair) x (array gene): a small set of query genes become hubs,
                   like the Kuzmin-2018 screen the reference was written for.

Nothing in this module touches the global ``random`` / ``numpy.random`` state (the reference's
``fold`` and ``initialize_parameters`` consume those streams and parity runs seed them).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "gene_names",
    "planted_triplets",
    "write_dat",
    "write_raw_s2",
    "planted_links_soa",
    "kuzmin_links_soa",
]


def gene_names(P: int) -> list[str]:
    """Gene names without '_' or tab (the fold-file format splits on both, TIP.py:331)."""
    return ["G%05d" % i for i in range(P)]


def _planted_model(rng: np.random.Generator, P: int, k_star: int):
    theta = rng.dirichlet(np.full(k_star, 0.3), size=P)
    p1 = rng.beta(0.5, 3.0, size=(k_star, k_star, k_star))
    return theta, p1


def _label(rng, theta, p1, a, b, c):
    prob = np.einsum("la,lb,lc,abc->l", theta[a], theta[b], theta[c], p1, optimize=True)
    return (rng.random(prob.shape[0]) < prob).astype(np.int8)


def planted_triplets(P: int, n_triplets: int, *, seed: int = 1, shape: str = "uniform",
                     k_star: int = 4, n_query: int | None = None):
    """Return (g[n,3] int32 sorted ascending per row, label[n] int8); rows are distinct triples.

    Every gene 0..P-1 is guaranteed to appear at least once (callers that fold the set must
    still check per-fold coverage, see ``ensure_train_coverage``)."""
    rng = np.random.default_rng(seed)
    theta, p1 = _planted_model(rng, P, k_star)
    seen: set[int] = set()
    out = np.empty((n_triplets, 3), dtype=np.int64)
    filled = 0
    # coverage chain first: (i, i+1, i+2) mod P for i = 0, 3, 6 ... touches every gene once
    chain = []
    for i in range(0, P, 3):
        t = sorted({i % P, (i + 1) % P, (i + 2) % P})
        if len(t) == 3:
            chain.append(t)
    chain = np.array(chain[:n_triplets], dtype=np.int64).reshape(-1, 3)
    P2 = P * P
    while filled < n_triplets:
        need = n_triplets - filled
        if filled == 0 and len(chain):
            cand = chain
        elif shape == "uniform":
            cand = rng.integers(0, P, size=(int(need * 1.2) + 16, 3))
        elif shape == "kuzmin":
            nq = n_query or max(4, int(round(P ** 0.5)))
            m = int(need * 1.2) + 16
            q = rng.integers(0, nq, size=(m, 2))
            arr = rng.integers(nq, P, size=(m, 1))
            cand = np.concatenate([q, arr], axis=1)
        else:
            raise ValueError("shape must be 'uniform' or 'kuzmin'")
        cand = np.sort(cand, axis=1)
        ok = (cand[:, 0] != cand[:, 1]) & (cand[:, 1] != cand[:, 2])
        cand = cand[ok]
        keys = cand[:, 0] * P2 + cand[:, 1] * P + cand[:, 2]
        # order-preserving de-duplication against everything emitted so far
        _, first = np.unique(keys, return_index=True)
        first.sort()
        for idx in first:
            k = int(keys[idx])
            if k in seen:
                continue
            seen.add(k)
            out[filled] = cand[idx]
            filled += 1
            if filled == n_triplets:
                break
    g = out.astype(np.int32)
    lab = _label(rng, theta, p1, g[:, 0], g[:, 1], g[:, 2])
    return g, lab


def write_dat(path: str, g: np.ndarray, lab: np.ndarray, names: list[str]) -> None:
    """Write the train/test format the reference reads (TIP.py:327-332): name_name_name<TAB>label."""
    with open(path, "w", encoding="utf-8") as fh:
        for (a, b, c), r in zip(g.tolist(), lab.tolist()):
            tri = sorted((names[a], names[b], names[c]))
            fh.write("_".join(tri) + "\t" + str(int(r)) + "\n")


def write_raw_s2(path: str, g: np.ndarray, lab: np.ndarray, names: list[str]) -> None:
    """Write an 8-column 'Data_S2'-style TSV that the reference's get_input digests (TIP.py:226-273).

    Columns: query strain, query alleles 'x+y', array strain, array allele, type, score, p-value, class.
    Label 1 is encoded as (score=-0.5, p=0.01), label 0 as (score=0.0, p=0.5)."""
    with open(path, "w", encoding="utf-8") as fh:
        fh.write("\t".join(["Query strain ID", "Query allele name", "Array strain ID", "Array allele name",
                            "Combined mutant type", "Adjusted genetic interaction score", "P-value",
                            "Interaction type"]) + "\n")
        for n, ((a, b, c), r) in enumerate(zip(g.tolist(), lab.tolist())):
            score, pval = ("-0.5", "0.01") if r else ("0.0", "0.5")
            fh.write("\t".join(["Q%d" % n, names[a] + "+" + names[b], "A%d" % n, names[c], "trigenic",
                                score, pval, "novel" if r else "none"]) + "\n")


def _decimal_string_rank(ids, xp):
    """Sort key that orders non-negative ids like Python orders their decimal strings (TIP.py:353 sorts the ids
    of a triplet AS STRINGS, so "10" < "9"): digits left-aligned to 10 places, shorter string first on a tie
    ("1" < "10").  Works on numpy arrays and torch tensors (int64)."""
    nd = xp.ones_like(ids)
    for d in range(1, 10):
        nd = nd + (ids >= 10 ** d)
    return (ids * 10 ** (10 - nd)) * 16 + nd


def kuzmin_links_soa(P: int, n_links: int, *, seed: int = 1, device=None, n_query: int | None = None):
    """Hub-shaped links at bench scale: (query pair) x (array gene) like the Kuzmin-2018 screen the reference
    digests (TIP.py:272-273), `n_query` (default round(sqrt(P)), 77 at P = 6000) query genes of degree
    ~2 n_links / n_query, every other gene an array gene.

    Ids are what the reference's digestion would give them: by first appearance (TIP.py:337-341), so the query
    genes - present in almost every line - hold the smallest ids and the array genes follow in random order; and
    the three ids of a link are put in the key's slot order, the DECIMAL-STRING order of TIP.py:353, so a hub
    lands in slot a, b or c exactly where it would in the reference ("1234" < "42" < "5").
    Returns int32 (g1, g2, g3, label) in slot order, torch tensors on `device` (numpy when device is None)."""
    nq = n_query or max(4, int(round(P ** 0.5)))
    if device is None:
        rng = np.random.default_rng(seed)
        q1 = rng.integers(0, nq, n_links)
        q2 = (q1 + 1 + rng.integers(0, nq - 1, n_links)) % nq
        arr = rng.integers(nq, P, n_links)
        ids = np.stack([q1, q2, arr], axis=1).astype(np.int64)
        order = np.argsort(_decimal_string_rank(ids, np), axis=1, kind="stable")
        ids = np.take_along_axis(ids, order, axis=1).astype(np.int32)
        lab = (rng.random(n_links) < 0.1).astype(np.int32)
        return ids[:, 0].copy(), ids[:, 1].copy(), ids[:, 2].copy(), lab
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    q1 = torch.randint(0, nq, (n_links,), device=device, generator=gen)
    q2 = (q1 + 1 + torch.randint(0, nq - 1, (n_links,), device=device, generator=gen)) % nq
    arr = torch.randint(nq, P, (n_links,), device=device, generator=gen)
    ids = torch.stack([q1, q2, arr], dim=1)
    order = torch.argsort(_decimal_string_rank(ids, torch), dim=1, stable=True)
    ids = torch.gather(ids, 1, order).to(torch.int32)
    lab = (torch.rand(n_links, device=device, generator=gen) < 0.1).to(torch.int32)
    return ids[:, 0].contiguous(), ids[:, 1].contiguous(), ids[:, 2].contiguous(), lab


def planted_links_soa(P: int, n_links: int, *, seed: int = 1, device=None, k_star: int = 4):
    """Large-scale generator for bench configs that cannot go through Python dicts (cfg4: 1e8 links).

    Returns int32 arrays (g1, g2, g3, label) of length n_links as torch tensors on ``device``
    (or numpy arrays when device is None).  Triples are drawn uniformly with distinct genes; at
    these densities (1e8 of 3.6e10 possible triples at P=6000) duplicates are a <0.3 % effect on
    throughput and are legal input (the reference counts them), so no rejection is done."""
    if device is None:
        rng = np.random.default_rng(seed)
        a = rng.integers(0, P, n_links, dtype=np.int32)
        b = (a + 1 + rng.integers(0, P - 1, n_links, dtype=np.int32)) % P
        c = rng.integers(0, P, n_links, dtype=np.int32)
        bad = (c == a) | (c == b)
        while bad.any():
            c[bad] = rng.integers(0, P, int(bad.sum()), dtype=np.int32)
            bad = (c == a) | (c == b)
        lab = (rng.random(n_links) < 0.1).astype(np.int32)
        return a, b.astype(np.int32), c, lab
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    a = torch.randint(0, P, (n_links,), device=device, dtype=torch.int32, generator=gen)
    b = (a + 1 + torch.randint(0, P - 1, (n_links,), device=device, dtype=torch.int32, generator=gen)) % P
    c = (b + 1 + torch.randint(0, P - 2, (n_links,), device=device, dtype=torch.int32, generator=gen)) % P
    c = torch.where(c == a, (c + 1) % P, c)
    c = torch.where(c == b, (c + 1) % P, c)
    lab = (torch.rand(n_links, device=device, generator=gen) < 0.1).to(torch.int32)
    return a, b.to(torch.int32), c.to(torch.int32), lab
