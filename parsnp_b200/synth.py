"""Synthetic genome generators for the parity tests and bench.py.

Definitions follow SURVEY.md section 8(d):
  G_indep(L, nq, div, seed): random reference, every query independently substituted at rate `div`
  G_pop(L, nq, div, seed):   2*div*L shared segregating sites, every query carries each alt allele with p=0.5
Genomes are returned as numpy uint8 arrays of ASCII codes ('A','C','G','T').
"""
import os
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def g_indep(L, nq, div, seed):
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, L).astype(np.uint8)
    out = [ref]
    for q in range(nq):
        r = np.random.default_rng(seed * 1000 + q + 1)
        mask = r.random(L) < div
        shift = r.integers(1, 4, L).astype(np.uint8)
        out.append(np.where(mask, (ref + shift) % 4, ref).astype(np.uint8))
    return [_ACGT[g] for g in out]


def g_pop(L, nq, div, seed):
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, L).astype(np.uint8)
    nseg = int(round(2 * div * L))
    seg = np.sort(rng.choice(L, nseg, replace=False))
    alt = ((ref[seg] + rng.integers(1, 4, nseg)) % 4).astype(np.uint8)
    out = [ref]
    for q in range(nq):
        r = np.random.default_rng([seed, q + 1])
        carry = r.random(nseg) < 0.5
        g = ref.copy()
        g[seg[carry]] = alt[carry]
        out.append(g)
    return [_ACGT[g] for g in out]


_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    _COMP[a] = b


def revcomp(g):
    return _COMP[g[::-1]]


def rearrange(g, rng, n_inv=2, inv_len=20000, dels=(37, 500), ins=(50,)):
    """optional realism add-ons of SURVEY 8(d): inversions, deletions, one insertion"""
    g = g.copy()
    L = len(g)
    for _ in range(n_inv):
        if L <= inv_len + 2:
            break
        s = int(rng.integers(0, L - inv_len))
        g[s:s + inv_len] = revcomp(g[s:s + inv_len])
    for d in dels:
        if len(g) <= d + 2:
            continue
        s = int(rng.integers(0, len(g) - d))
        g = np.concatenate([g[:s], g[s + d:]])
    for i in ins:
        s = int(rng.integers(0, len(g)))
        g = np.concatenate([g[:s], _ACGT[rng.integers(0, 4, i)], g[s:]])
    return g


def with_contig_padding(g, contigs, d=300):
    """a multi-contig QUERY as parsnp_core ingests it: d+10 N's at every contig break (src/parsnp.cpp:3114-3118); the reference
    gets none. Same contig bounds as write_fasta(..., contigs=contigs)."""
    L = len(g)
    bounds = [L * i // contigs for i in range(contigs + 1)]
    pad = np.full(d + 10, ord("N"), np.uint8)
    parts = []
    for c in range(contigs):
        if c:
            parts.append(pad)
        parts.append(g[bounds[c]:bounds[c + 1]])
    return np.concatenate(parts)


def write_fasta(path, name, seq, width=80, contigs=1):
    """seq: uint8 ASCII array. contigs>1 splits into equal contigs with their own headers."""
    L = len(seq)
    with open(path, "wb") as f:
        bounds = [L * i // contigs for i in range(contigs + 1)]
        for c in range(contigs):
            hdr = name if contigs == 1 else "%s_c%d" % (name, c + 1)
            f.write(b">" + hdr.encode() + b"\n")
            s = seq[bounds[c]:bounds[c + 1]]
            n = len(s)
            full = (n // width) * width
            if full:
                body = np.empty((n // width, width + 1), dtype=np.uint8)
                body[:, :width] = s[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")


def write_dataset(outdir, genomes, contigs=1):
    """writes ref.fna + q%04d.fna, returns (ref_path, [query paths])"""
    os.makedirs(outdir, exist_ok=True)
    ref = os.path.join(outdir, "ref.fna")
    write_fasta(ref, "ref", genomes[0], contigs=contigs)
    qs = []
    for k, g in enumerate(genomes[1:]):
        p = os.path.join(outdir, "q%04d.fna" % k)
        write_fasta(p, "q%d" % k, g, contigs=contigs)
        qs.append(p)
    return ref, qs
