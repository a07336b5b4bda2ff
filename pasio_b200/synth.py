"""Seeded synthetic coverage profiles for the five BASELINE.json configs.

Host-side numpy generators only (inputs for tests, bench.py and the golden
script); nothing here is on the segmentation path.  Definitions follow
SURVEY.md section 8(d).
"""
import numpy as np

# hg38 contig-size profile used as the *synthetic profile definition* of config 4
HG38_SIZES = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555),
    ("chr5", 181538259), ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636),
    ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
    ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345),
    ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415),
]


def piecewise_poisson(n, seed):
    """Config 1/3 generator: segments L=int(Exp(500))+1, lambda~Gamma(1,10), Poisson(lambda)."""
    rs = np.random.RandomState(seed)
    parts, total = [], 0
    while total < n:
        seg_len = int(rs.exponential(500)) + 1
        lam = rs.gamma(1.0, 10.0)
        parts.append(rs.poisson(lam, seg_len))
        total += seg_len
    return np.concatenate(parts)[:n].astype(np.int64)


def two_level_poisson(half, seed=2, lam_a=15, lam_b=20):
    """tests/test_bench_pasio.py:42-43 style profile: Poisson(15) half then Poisson(20) half."""
    rs = np.random.RandomState(seed)
    return np.concatenate([rs.poisson(lam_a, half), rs.poisson(lam_b, half)]).astype(np.int64)


def dnase_like(n, seed, hotspot_share=0.10):
    """Config 2/4/5 generator: sparse background blocks alternating with hotspots.

    background: L=Exp(20000)+1, lambda~Gamma(0.5, 0.04); hotspot: L=Exp(400)+50,
    lambda~Gamma(2, 3).  Built block-wise so a 248 Mb contig takes seconds.
    """
    rs = np.random.RandomState(seed)
    out = np.zeros(n, dtype=np.int64)
    pos = 0
    while pos < n:
        if rs.random_sample() < hotspot_share:
            seg_len = int(rs.exponential(400)) + 50
            lam = rs.gamma(2.0, 3.0)
        else:
            seg_len = int(rs.exponential(20000)) + 1
            lam = rs.gamma(0.5, 0.04)
        seg_len = min(seg_len, n - pos)
        out[pos:pos + seg_len] = rs.poisson(lam, seg_len)
        pos += seg_len
    return out


def random_candidates(n, n_candidates, seed):
    """Config 3: {0,n} plus sorted distinct random interior positions."""
    rs = np.random.RandomState(seed)
    inner = rs.choice(np.arange(1, n), size=n_candidates - 2, replace=False)
    return np.concatenate([[0], np.sort(inner), [n]]).astype(np.int64)


def genome_profile(scale=1.0, n_scaffolds=170, seed=4):
    """Config 4 contig lengths: 24 chromosomes + chrM + log-uniform scaffolds, scaled."""
    rs = np.random.RandomState(seed)
    sizes = [(name, max(1000, int(size * scale))) for name, size in HG38_SIZES]
    sizes.append(("chrM", 16569))
    scaf = np.exp(rs.uniform(np.log(1e3), np.log(4.5e5), n_scaffolds)).astype(np.int64)
    sizes.extend(("scaffold%d" % i, int(s)) for i, s in enumerate(scaf))
    return sizes


def transcript_lengths(n_contigs, seed=5):
    """Config 5 contig lengths: uniform int [1000, 50000]."""
    rs = np.random.RandomState(seed)
    return rs.randint(1000, 50001, size=n_contigs).astype(np.int64)


def to_bedgraph_lines(chrom, counts, chrom_start=0):
    """Run-length encode a dense profile as bedgraph text lines."""
    counts = np.asarray(counts)
    change = np.flatnonzero(counts[1:] != counts[:-1]) + 1
    starts = np.concatenate([[0], change])
    stops = np.concatenate([change, [len(counts)]])
    vals = counts[starts]
    return ["%s\t%d\t%d\t%d\n" % (chrom, s + chrom_start, e + chrom_start, v)
            for s, e, v in zip(starts.tolist(), stops.tolist(), vals.tolist())]
