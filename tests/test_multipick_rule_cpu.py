"""CPU suite: the ACCEPTANCE RULE of the multi-pick FPS rounds (csrc/fps_bucket.cu, K > 1), restated in numpy and checked
against the oracle's one-pick-at-a-time FPS.  Integer coordinates make every squared distance exact in float32, so the
emulation needs no FMA modelling and the clouds are full of exact ties -- the hard part of the rule: a later candidate of a
round may only be taken when it is provably the pick the sequential algorithm makes next.

Rule under test: the points are split into groups ("warps"); each group posts its candidate (largest
min-distance, smallest reference rank among equals) and a runner-up key (largest min-distance among its other points; exact
duplicates of the candidate excepted).  Position 0 of the key-sorted candidates is the plain argmax.  Position k > 0 is
accepted while (i) no other candidate has the same key (the kernel: within 32 ulp), (ii) its key is strictly above the
runner-up of every group an earlier pick of the round came from, (iii) no earlier pick of the round lowers it."""
import numpy as np
import pytest


def _ref_rank(k, bs):
    L = bs.bit_length() - 1
    low = k & (bs - 1)
    rev = int(format(low, f"0{L}b")[::-1], 2) if L else 0
    return (rev << 32) | (k >> L)  # bit-reversed thread slot first, then the stride count (oracle/pointnet2_oracle.c)


def _multipick_fps(xyz, m, groups, K, bs):
    n = len(xyz)
    order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))  # any spatially sorted assignment of points to groups
    gid = np.empty(n, np.int64)
    gid[order] = np.arange(n) * groups // n
    rank = np.array([_ref_rank(k, bs) for k in range(n)], dtype=np.int64)
    md = np.full(n, 1e10, np.float32)
    picks, pending, rounds = [0], [0], 0
    while len(picks) < m:
        for c in pending:
            d = ((xyz - xyz[c]) ** 2).sum(1).astype(np.float32)
            md = np.minimum(md, d)
        rounds += 1
        recs = []  # (key, best point, runner-up key)
        for g in range(groups):
            idx = np.nonzero(gid == g)[0]
            if len(idx) == 0:
                continue
            top = md[idx].max()
            tied = idx[md[idx] == top]
            best = tied[np.argmin(rank[tied])]
            others = idx[idx != best]
            if len(tied) > 1 and (xyz[tied] == xyz[best]).all():  # only exact duplicates share the maximum
                others = idx[md[idx] != top]
            run = md[others].max() if len(others) else -1.0
            recs.append((float(top), int(best), float(run)))
        recs.sort(key=lambda r: (-r[0], rank[r[1]]))  # position 0 = largest key, smallest reference rank
        pending, run2 = [recs[0][1]], recs[0][2]
        for k in range(1, min(K, len(recs))):
            key, q, run = recs[k]
            if len(picks) + len(pending) >= m:
                break
            shared = key == recs[k - 1][0] or (k + 1 < len(recs) and key == recs[k + 1][0])
            moved = any(np.float32(((xyz[p] - xyz[q]) ** 2).sum()) < np.float32(key) for p in pending)
            if shared or not key > run2 or not key > 0.0 or moved:
                break
            pending.append(q)
            run2 = max(run2, run)
        picks += pending
    return np.array(picks[:m], np.int32), rounds


def _clouds():
    rng = np.random.default_rng(0)
    lattice = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
    rng.shuffle(lattice)
    uniq = rng.integers(0, 200, size=(700, 3))
    dup = np.concatenate([uniq, uniq[rng.integers(0, 700, size=324)]])
    rng.shuffle(dup)
    sparse = rng.integers(0, 1000, size=(1024, 3))
    few = np.repeat(rng.integers(0, 50, size=(40, 3)), 16, axis=0)
    rng.shuffle(few)
    return {"lattice": lattice, "dup": dup, "sparse": sparse, "few_unique": few}


@pytest.mark.parametrize("name", ["lattice", "dup", "sparse", "few_unique"])
@pytest.mark.parametrize("K", [2, 8])
def test_multi_pick_rule_reproduces_sequential_fps(name, K):
    from oracle import oracle as orc

    pts = _clouds()[name].astype(np.float32)
    n = len(pts)
    m = min(300, n)
    want = orc.fps(pts[None], m)[0]
    got, rounds = _multipick_fps(pts, m, groups=8, K=K, bs=orc.opt_n_threads(n))
    assert np.array_equal(got, want), f"first mismatch at pick {np.argwhere(got != want)[:3].ravel()}"
    if name == "sparse":
        assert rounds < 0.8 * m  # the rule does accept several picks per round on a generic cloud


# ---------------------------------------------------------------------------------------------------------------------
# The same rule over a CLUSTER (csrc/fps_bucket_cluster.cu, fps_bucket_cluster_mp_kernel): the groups belong to CTAs; a CTA
# only publishes its top-KX candidates (entry 0 = its exact argmax; cut to that one entry when the top is shared) and
# `ubound`, the largest key it did not publish.  A merged position k > 0 additionally needs its key strictly above every
# CTA's ubound -- every unpublished candidate is then strictly below it.
def _multipick_fps_cluster(xyz, m, ctas, groups_per_cta, KX, KA, bs, drop_ubound=False):
    n = len(xyz)
    order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
    groups = ctas * groups_per_cta
    gid = np.empty(n, np.int64)
    gid[order] = np.arange(n) * groups // n
    rank = np.array([_ref_rank(k, bs) for k in range(n)], dtype=np.int64)
    md = np.full(n, 1e10, np.float32)
    picks, pending, rounds = [0], [0], 0
    while len(picks) < m:
        for c in pending:
            md = np.minimum(md, ((xyz - xyz[c]) ** 2).sum(1).astype(np.float32))
        rounds += 1
        listed, umax = [], -1.0
        for cta in range(ctas):
            recs = []
            for g in range(cta * groups_per_cta, (cta + 1) * groups_per_cta):
                idx = np.nonzero(gid == g)[0]
                if len(idx) == 0:
                    continue
                top = md[idx].max()
                tied = idx[md[idx] == top]
                best = tied[np.argmin(rank[tied])]
                others = idx[idx != best]
                if len(tied) > 1 and (xyz[tied] == xyz[best]).all():
                    others = idx[md[idx] != top]
                run = md[others].max() if len(others) else -1.0
                recs.append((float(top), int(best), float(run)))
            if not recs:
                continue
            recs.sort(key=lambda r: (-r[0], rank[r[1]]))
            if len(recs) > 1 and recs[1][0] == recs[0][0]:  # shared top: only the exact local argmax is published
                listed.append(recs[0])
                umax = max(umax, recs[0][0])
            else:
                listed += recs[:KX]
                if len(recs) > KX:
                    umax = max(umax, recs[KX][0])
        if drop_ubound:
            umax = -1.0
        listed.sort(key=lambda r: (-r[0], rank[r[1]]))
        pending, run2 = [listed[0][1]], listed[0][2]
        top_shared = len(listed) > 1 and listed[1][0] == listed[0][0]
        for k in range(1, min(KA, len(listed))):
            key, q, run = listed[k]
            if len(picks) + len(pending) >= m:
                break
            shared = top_shared or key == listed[k - 1][0] or (k + 1 < len(listed) and key == listed[k + 1][0])
            moved = any(np.float32(((xyz[p] - xyz[q]) ** 2).sum()) < np.float32(key) for p in pending)
            if shared or not key > umax or not key > run2 or not key > 0.0 or moved:
                break
            pending.append(q)
            run2 = max(run2, run)
        picks += pending
    return np.array(picks[:m], np.int32), rounds


@pytest.mark.parametrize("name", ["lattice", "dup", "sparse", "few_unique"])
@pytest.mark.parametrize("ctas,KX", [(2, 4), (5, 2), (3, 1)])
def test_cluster_multi_pick_rule_reproduces_sequential_fps(name, ctas, KX):
    from oracle import oracle as orc

    pts = _clouds()[name].astype(np.float32)
    n = len(pts)
    m = min(300, n)
    want = orc.fps(pts[None], m)[0]
    got, rounds = _multipick_fps_cluster(pts, m, ctas=ctas, groups_per_cta=4, KX=KX, KA=8, bs=orc.opt_n_threads(n))
    assert np.array_equal(got, want), f"first mismatch at pick {np.argwhere(got != want)[:3].ravel()}"
    if name == "sparse" and KX > 1:
        assert rounds < 0.8 * m


def test_cluster_rule_needs_the_unlisted_bound():
    """Dropping the `ubound` test (a candidate above every PUBLISHED key may still be below an unpublished one) breaks the
    result -- the test is not vacuous."""
    from oracle import oracle as orc

    pts = _clouds()["sparse"].astype(np.float32)
    n, m = len(pts), 300
    want = orc.fps(pts[None], m)[0]
    got, _ = _multipick_fps_cluster(pts, m, ctas=2, groups_per_cta=8, KX=1, KA=8, bs=orc.opt_n_threads(n), drop_ubound=True)
    assert not np.array_equal(got, want)
