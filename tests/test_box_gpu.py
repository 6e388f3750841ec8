"""Parity of the box-kernel confidence path (RLS.py:120-181) with the logged field queries and
the oracle: visit counts and the picked action bit-exact, mean/var within 1e-5."""
import os

import numpy as np
import pytest
import torch

import dcarl_b200 as dc
from dcarl_b200 import _abi
from oracle import cport, pyref

pytestmark = pytest.mark.gpu
DIST = dc.VISITED_STATE_DIST[0]


def _f32(x):
    return np.asarray(x, np.float32).astype(np.float64)


def _oracle_nmv(vs, ret, queries, n_actions, dist=DIST):
    """oracle on the SAME inputs the device holds (float32 buffer widened to float64)."""
    return cport.box_moments(_f32(vs), _f32(ret), dist, queries, n_actions)


def _check(buf, vs, ret, queries, n_actions, rls=None, dist=DIST):
    n, mean, var, pick = buf.query(queries, n_actions, rls)
    ref = _oracle_nmv(vs, ret, queries, n_actions, dist)
    assert np.array_equal(n, ref[:, :, 0])                                  # visited times: exact
    assert np.abs(mean - ref[:, :, 1]).max() < 1e-9 and np.abs(var - ref[:, :, 2]).max() < 1e-9
    kw = {} if rls is None else dict(visited_times_thres=rls.visited_times_thres, confidence_thres=rls.confidence_thres)
    assert np.array_equal(pick, cport.box_select(ref, **kw))
    return n, mean, var, pick


def test_logged_field_queries(golden):
    g = golden("rls_box")
    vs, vv, dr = g["visited_state"], g["visited_value"], g["driving_record"]
    buf = dc.BoxBuffer(20)
    buf.append(vs, vv[:, 1])
    n, mean, var, pick = _check(buf, vs, vv[:, 1], dr[:, :20], 8)
    assert np.array_equal(n[:, 0], dr[:, 23])                               # 107 / 107 logged counts
    assert np.abs(mean[:, 0] - dr[:, 24]).max() < 1e-5 and np.abs(var[:, 0] - dr[:, 25]).max() < 1e-5
    # one query at a time gives the same answer as the batch
    n1, m1, v1, p1 = buf.query(dr[17, :20], 8)
    assert np.array_equal(n1[0], n[17]) and p1[0] == pick[17]


def _clustered(rng, n_centers=6, n_actions=8, max_rows=40, D=20):
    centers = rng.normal(0, 1, (n_centers, D)) * DIST[:D]
    rows, vals = [], []
    for c in centers:
        for a in range(n_actions):
            m = int(rng.integers(0, max_rows))
            s = c + rng.uniform(-0.45, 0.45, (m, D)) * DIST[:D]
            rows.append(np.hstack([s, np.full((m, 1), float(a))]))
            vals.append(-0.95 ** rng.integers(0, 10, m) * (rng.random(m) < (0.3 + 0.08 * a)))
    return centers, np.vstack(rows), np.concatenate(vals)


def test_candidate_branch_and_literal_port():
    rng = np.random.default_rng(11)
    centers, vs, ret = _clustered(rng)
    buf = dc.BoxBuffer(20)
    buf.append(vs, ret)
    n, mean, var, pick = _check(buf, vs, ret, centers, 8)
    assert pick.max() > 0
    lit = pyref.BoxBuffer(_f32(vs), np.stack([vs[:, 20], _f32(ret)], 1))
    for qi, c in enumerate(centers):
        assert int(lit.act_test(c)) == pick[qi]
    # a stricter confidence threshold goes through Phi(z)
    rp = _abi.RlsParams(confidence_thres=0.9, visited_times_thres=3)
    _check(buf, vs, ret, centers, 8, rp)


@pytest.mark.parametrize("M", [0, 1, 255, 256, 257, 1000, 4099])
def test_ragged_buffer_sizes(M):
    rng = np.random.default_rng(M)
    _, vs, ret = _clustered(rng, n_centers=3, max_rows=200)
    assert len(vs) >= M or M == 4099
    reps = -(-max(M, 1) // len(vs))
    vs, ret = np.tile(vs, (reps, 1))[:M], np.tile(ret, reps)[:M]
    buf = dc.BoxBuffer(20, capacity=16)
    if M:
        buf.append(vs, ret)
    q = vs[rng.integers(0, max(M, 1), 5), :20] if M else np.zeros((5, 20))
    _check(buf, vs.reshape(-1, 21), ret, q, 8)


def test_closed_box_faces_and_float32_threshold_walk():
    """Stored rows exactly on, just inside and just outside a face of the query box."""
    rng = np.random.default_rng(5)
    q = (rng.normal(0, 3, 20)).astype(np.float32).astype(np.float64)
    rows = []
    for d in range(20):
        for sign in (-1.0, 1.0):
            face = np.float32(q[d] + sign * DIST[d])
            for s in (face, np.nextafter(face, np.float32(np.inf)), np.nextafter(face, np.float32(-np.inf))):
                r = q.copy()
                r[d] = s
                rows.append(np.append(r, 0.0))
    vs = np.asarray(rows)
    ret = -rng.random(len(vs))
    buf = dc.BoxBuffer(20)
    buf.append(vs, ret)
    n, *_ = _check(buf, vs, ret, q[None], 4)
    assert 0 < n[0, 0] < len(vs)


def test_many_queries_many_candidates_and_non_integer_actions():
    rng = np.random.default_rng(8)
    M, Q, NA = 20000, 64, 256
    base = rng.normal(0, 1, (40, 20)) * DIST[:20]
    vs = base[rng.integers(0, 40, M)] + rng.normal(0, 0.3, (M, 20)) * DIST[:20]
    acts = rng.integers(0, NA, M).astype(np.float64)
    acts[::97] += 0.05                 # inside the +-0.1 action tolerance
    acts[::89] += 0.3                  # outside: matches no candidate
    vs = np.hstack([vs, acts[:, None]])
    ret = -rng.random(M)
    buf = dc.BoxBuffer(20)
    buf.append(vs, ret)
    queries = base[rng.integers(0, 40, Q)] + rng.normal(0, 0.1, (Q, 20)) * DIST[:20]
    n, *_ = _check(buf, vs, ret, queries, NA)
    assert n.sum() > 0
    # a wide action tolerance lets one stored row count for several candidates (general rtree semantics)
    wide = DIST.copy()
    wide[20] = 1.0
    buf2 = dc.BoxBuffer(20, dist=wide)
    buf2.append(vs[:3000], ret[:3000])
    _check(buf2, vs[:3000], ret[:3000], queries[:4], 16, dist=wide)


def test_other_dimensions_use_the_generic_kernel():
    rng = np.random.default_rng(13)
    D = 6
    dist = np.array([1, 0.5, 2, 3, 0.25, 4, 0.1])
    vs = np.hstack([rng.normal(0, 1, (5000, D)), rng.integers(0, 4, (5000, 1)).astype(np.float64)])
    ret = rng.normal(0, 1, 5000)
    buf = dc.BoxBuffer(D, dist=dist)
    buf.append(vs, ret)
    q = rng.normal(0, 0.5, (9, D))
    n, mean, var, _ = buf.query(q, 4)
    ref = cport.box_moments(_f32(vs), _f32(ret), dist, q, 4)
    assert np.array_equal(n, ref[:, :, 0]) and np.abs(mean - ref[:, :, 1]).max() < 1e-9
    assert np.abs(var - ref[:, :, 2]).max() < 1e-9


def test_rls_class_drop_in(tmp_path, golden, monkeypatch):
    """Constructor loads visited_*.txt from the cwd, act/act_test answer like the logged run,
    add_data appends lagged / terminal-discounted returns (RLS.py:185-215)."""
    g = golden("rls_box")
    vs, vv, dr = g["visited_state"], g["visited_value"], g["driving_record"]
    monkeypatch.chdir(tmp_path)
    np.savetxt("visited_state.txt", vs, fmt="%f")
    np.savetxt("visited_value.txt", vv, fmt="%f")
    rls = dc.RLS(visited_times_thres=10, is_training=False, debug=False, save_new_data=True,
                 create_new_train_file=False, create_new_record_file=True, save_new_driving_data=True)
    assert rls.visited_state_counter == len(vs)
    for qi in (0, 50, 106):
        obs = dr[qi, :20]
        rule = rls.state_with_action(obs, 0)
        assert rls._calculate_visited_times(rule, rls.visited_state_tree) == int(dr[qi, 23])
        mean, var, sigma = rls._calculate_statistics_index(rule, rls.visited_state_value, rls.visited_state_tree)
        assert abs(mean - dr[qi, 24]) < 1e-5 and abs(var - dr[qi, 25]) < 1e-5 and abs(sigma - np.sqrt(var)) < 1e-12
        a = rls.act(obs, np.array(3))
        assert isinstance(a, np.ndarray) and a.shape == () and int(a) == 0
    # unseen observation: the sentinel
    far = np.full(20, 1e4)
    assert rls._calculate_statistics_index(rls.state_with_action(far, 2)) == (-1, -1, -1)
    # add_data: 14 quiet steps then a collision; 15 new rows with the discounted tail
    before = len(rls.visited_state_tree)
    for t in range(15):
        rls.add_data(far + t, 1, 0.0 if t < 14 else -1.0, far + t + 1, t == 14)
    assert len(rls.visited_state_tree) == before + 15
    tail = rls.visited_state_tree.ret[before:].cpu().numpy()
    want = np.array([r for _, r in pyref.add_data_returns([0.0] * 14 + [-1.0], [False] * 14 + [True])], np.float64)
    assert tail.dtype == np.float64 and np.array_equal(tail, want)          # the drop-in stores float64 like the reference
    rls.visited_state_outfile.flush(); rls.driving_record_outfile.flush()
    assert len(open("visited_state.txt").readlines()) == len(vs) + 15
    rec = np.loadtxt("driving_record.txt")
    assert rec.shape == (15, 29)
    # should_use_rule: under-visited rule action -> rule
    rls.is_training = True
    assert int(rls.act(far, np.array(5))) == 0


def test_training_mode_gate_follows_the_reference_draw_for_draw(tmp_path, golden, monkeypatch):
    """should_use_rule / act_train (RLS.py:84-118) with a seeded `random`: the under-visited branch, and the
    exploration branch `random.uniform(-1, 0) < mean_rule` on both of its outcomes, against the literal port
    consuming the same random stream."""
    import random
    g = golden("rls_box")
    vs, vv, dr = g["visited_state"], g["visited_value"], g["driving_record"]
    monkeypatch.chdir(tmp_path)
    np.savetxt("visited_state.txt", vs, fmt="%f")
    np.savetxt("visited_value.txt", vv, fmt="%f")
    rls = dc.RLS(visited_times_thres=10, is_training=True, debug=False, save_new_data=False,
                 create_new_train_file=False, create_new_record_file=True, save_new_driving_data=False)
    lit = pyref.BoxBuffer(np.loadtxt("visited_state.txt"), np.loadtxt("visited_value.txt"))
    # stored observations whose rule action is well visited and whose mean return lies inside (-1, 0)
    rows = np.nonzero((vs[:, 20] == 0) & (vv[:, 1] < -0.3))[0]               # neighbourhoods with negative returns
    raw = cport.box_raw(vs, vv[:, 1], g["dist"].ravel(), vs[rows, :20], 1, threads=4)
    mean = raw[:, 0, 1] / np.maximum(raw[:, 0, 0], 1)
    cand = [int(r) for r, n, m in zip(rows, raw[:, 0, 0], mean) if n >= 10 and -0.9 < m < -0.1]
    assert len(cand) >= 5
    dr = vs                                   # (the loop below indexes observations by row)
    outcomes = set()
    for seed in range(6):
        for i in cand[:12]:
            obs = dr[i, :20]
            random.seed(1000 * seed + i)
            want = bool(lit.should_use_rule(obs, visited_times_thres=10))
            state_after = random.getstate()
            random.seed(1000 * seed + i)
            got = rls.should_use_rule(obs)
            assert got == want and random.getstate() == state_after          # same decision, same number of draws
            random.seed(1000 * seed + i)
            a = rls.act(obs, np.array(4))                                     # is_training -> act_train
            assert int(a) == (0 if want else 4)
            outcomes.add(want)
    assert outcomes == {True, False}
    far = np.full(20, 1e4)                                                    # under-visited: rule, and no draw at all
    random.seed(5)
    s0 = random.getstate()
    assert rls.should_use_rule(far) is True and random.getstate() == s0


def test_query_graph_replay_matches_the_plain_path_and_tracks_appends():
    rng = np.random.default_rng(21)
    centers, vs, ret = _clustered(rng, n_centers=5)
    buf = dc.BoxBuffer(20)
    half = len(vs) // 2
    buf.append(vs[:half], ret[:half])
    run = buf.query_graph(len(centers), 8)
    for upto in (half, len(vs)):
        mom, mv, pick = run(centers)
        torch.cuda.synchronize()
        n, mean, var, p = buf.query(centers, 8)
        assert np.array_equal(mom[:, :, 0].cpu().numpy(), n) and np.array_equal(pick.cpu().numpy(), p)
        assert np.allclose(mv[:, :, 0].cpu().numpy(), mean, rtol=0, atol=1e-12)
        ref = _oracle_nmv(vs[:upto], ret[:upto], centers, 8)
        assert np.array_equal(n, ref[:, :, 0])
        if upto == half:
            buf.append(vs[half:], ret[half:])          # the next call must re-capture with the new row count


def test_query_graph_is_not_disturbed_by_other_calls_on_the_buffer():
    """The captured launch owns its workspace: a larger moments() call between two replays (which replaces
    the buffer's shared threshold workspace) or a second graph must not change what a replay returns."""
    rng = np.random.default_rng(22)
    centers, vs, ret = _clustered(rng, n_centers=6)
    buf = dc.BoxBuffer(20)
    buf.append(vs, ret)
    run = buf.query_graph(2, 8)
    want = _oracle_nmv(vs, ret, centers[:2], 8)
    mom, _, pick = run(centers[:2])
    first = (mom.clone(), pick.clone())
    big = np.repeat(centers, 40, axis=0)                                 # Q = 240: a much larger workspace
    buf.moments(big, 8)
    other = buf.query_graph(6, 8)
    other(centers)
    junk = torch.full((1 << 20,), 0x7f, dtype=torch.uint8, device="cuda")  # reuse whatever the allocator freed
    mom, _, pick = run(centers[:2])
    torch.cuda.synchronize()
    assert torch.equal(mom, first[0]) and torch.equal(pick, first[1])
    assert np.array_equal(mom[:, :, 0].cpu().numpy(), want[:, :, 0])
    del junk


def test_fused_query_launch_matches_the_three_step_path():
    """dcarl_box_query (thresholds + scan + totals + select in one launch, two alternating accumulators) against
    dcarl_box_moments + dcarl_box_select, including the global-atomics bins, chunked queries, repeated
    calls on one workspace and an empty buffer."""
    from dcarl_b200 import dist as dd
    rng = np.random.default_rng(23)
    centers, vs, ret = _clustered(rng, n_centers=40, n_actions=12, max_rows=30)
    buf = dc.BoxBuffer(20)
    buf.append(vs, ret)
    for Q, na in ((1, 8), (7, 12), (40, 12), (40, 256)):                 # 40 x 256 bins: global atomics; > smem: chunks
        q = torch.from_numpy(np.ascontiguousarray(centers[:Q])).cuda()
        mom = buf.moments(q, na)
        mv, pick = buf.select(mom)
        fq = dd.ShardedBoxQuery(buf, Q, na)
        for rep in range(3):
            fmom, fmv, fpick = fq(q)
            torch.cuda.synchronize()
            assert torch.equal(fmom[:, :, 0], mom[:, :, 0]) and torch.equal(fpick, pick), (Q, na, rep)
            assert float((fmom - mom).abs().max()) < 1e-9 and float((fmv - mv).abs().max()) < 1e-9
        off = (-fq.ws.data_ptr()) % 16                                      # header {ticket, calls, extent} of the workspace
        hdr = fq.ws[off:off + 12].view(torch.int32).tolist()
        assert hdr == [0, 3, Q * na * 3]                                    # ticket back at zero, three calls counted
        ref = _oracle_nmv(vs, ret, centers[:Q], na)
        assert np.array_equal(fmom[:, :, 0].cpu().numpy(), ref[:, :, 0])
    empty = dc.BoxBuffer(20)
    fq = dd.ShardedBoxQuery(empty, 3, 8)
    fmom, fmv, fpick = fq(centers[:3])
    torch.cuda.synchronize()
    assert float(fmom.abs().sum()) == 0 and bool((fmv == -1).all()) and bool((fpick == 0).all())


def test_one_workspace_serves_queries_of_changing_shape():
    """BoxBuffer.query keeps ONE workspace for the one-launch query path: the two accumulators alternate between
    calls and a call zeroes what the previous call used of the other one, whatever the shapes were (more queries
    than one launch takes, one query, more candidate actions, back to many)."""
    rng = np.random.default_rng(41)
    centers, vs, ret = _clustered(rng, n_centers=80, n_actions=12, max_rows=25)
    buf = dc.BoxBuffer(20)
    buf.append(vs, ret)
    for Q, na in ((70, 8), (1, 8), (3, 40), (70, 12), (2, 256), (70, 8), (1, 12)):
        q = centers[:Q] + rng.uniform(-0.2, 0.2, (Q, 20)) * DIST[:20]
        n, mean, var, pick = buf.query(q, na)
        ref = _oracle_nmv(vs, ret, q, na)
        assert np.array_equal(n, ref[:, :, 0]), (Q, na)
        assert np.abs(mean - ref[:, :, 1]).max() < 1e-9 and np.abs(var - ref[:, :, 2]).max() < 1e-9
        assert np.array_equal(pick, cport.box_select(ref)), (Q, na)


def test_stats_chain_of_a_mixed_batch():
    """dcarl_stream_stats_reduce: zero-padded width and accumulation across the shapes of a mixed batch."""
    from dcarl_b200 import dist as dd
    from helpers import synth_streams
    i1, a1, R1 = synth_streams(24, 300, 1, 30, seed=1, r_dtype=np.float32, live_actions=11)
    i2, a2, R2 = synth_streams(24, 300, 20, 11, seed=2, r_dtype=np.float32)
    p1, p2 = dc.StreamPlan(24, 300, 1, 30), dc.StreamPlan(24, 300, 20, 11)
    d1 = [torch.from_numpy(x).cuda() for x in (i1, a1, R1)]
    d2 = [torch.from_numpy(x).cuda() for x in (i2, a2, R2)]
    want = dc.merge_stats([p1.run_device(*d1).stats.clone(), p2.run_device(*d2).stats.clone()])
    for keep in (True, False):
        mb = dd.MixedStreamBatch([p1, p2], keep_local=keep)
        got = mb.run([d1, d2]).clone()
        got2 = mb.run([d1, d2])
        torch.cuda.synchronize()
        assert got.numel() == 35 and torch.equal(got, want) and torch.equal(got2, want)


def test_float64_storage_is_exact_on_box_faces_and_float32_is_its_own_oracle(golden):
    """Queries within +-1e-5 of box faces of the shipped buffer: the float64 device buffer reproduces the
    float64 oracle count for count (the reference's precision, RLS.py:195-196); the float32 buffer reproduces
    the oracle on the narrowed rows, and differs from the float64 one on a fraction of a percent of them."""
    from helpers import face_queries
    g = golden("rls_box")
    vs, vv, dist = g["visited_state"], g["visited_value"], g["dist"].ravel()
    q, _ = face_queries(vs, dist, 6000, seed=9)
    want64 = cport.box_raw(vs, vv[:, 1], dist, q, 8, threads=4)
    want32 = cport.box_raw(vs.astype(np.float32), vv[:, 1].astype(np.float32), dist, q, 8, threads=4)
    b64 = dc.BoxBuffer(20, dtype=torch.float64)
    b64.append(vs, vv[:, 1])
    b32 = dc.BoxBuffer(20)
    b32.append(vs, vv[:, 1])
    from dcarl_b200 import dist as dd
    for buf, want in ((b64, want64), (b32, want32)):
        mom = buf.moments(q, 8).cpu().numpy()
        assert np.array_equal(mom[:, :, 0], want[:, :, 0])
        assert np.abs(mom[:, :, 1:] - want[:, :, 1:]).max() < 1e-9
        fq = dd.ShardedBoxQuery(buf, 64, 8)                                  # the fused launch, same answer
        fm, _, _ = fq(q[:64])
        assert np.array_equal(fm[:, :, 0].cpu().numpy(), want[:64, :, 0])
    assert (want64[:, :, 0] != want32[:, :, 0]).any()
    # the drop-in class stores float64
    n64, mean64, var64, pick64 = b64.query(g["driving_record"][:, :20], 8)
    assert np.array_equal(n64[:, 0], g["driving_record"][:, 23])
    # the sorted index and appended tails work in float64 too
    b64.build_index()
    assert np.array_equal(b64.moments_indexed(q[:5], 8).cpu().numpy()[:, :, 0], want64[:5, :, 0])


def test_many_queries_take_the_signature_filter_and_agree_with_the_plain_scan():
    """Q >= 8 on a float32 buffer of 20 dims: per-dimension bitmap pre-filter + exact test on the survivors.  Counts
    must equal the oracle's for clustered data, for queries far outside the data, for duplicate queries, NaN
    queries and rows on the edge of the queries' range."""
    rng = np.random.default_rng(77)
    centers, vs, ret = _clustered(rng, n_centers=70, n_actions=9, max_rows=25)
    vs = np.vstack([vs, np.hstack([centers[:30] + DIST[:20], np.zeros((30, 1))])])     # rows exactly on a box face
    ret = np.concatenate([ret, -np.ones(30)])
    buf = dc.BoxBuffer(20)
    buf.append(vs, ret)
    q = np.vstack([centers[:60], centers[:2], centers[5:7] + 1e3, centers[:1] * np.nan, centers[3:4] - 0.999 * DIST[:20]])
    assert len(q) == 66                                                                 # > 64: two filtered launches
    for Q in (8, 33, 64, 66):
        mom = buf.moments(q[:Q], 9).cpu().numpy()
        want = _oracle_nmv(vs, ret, q[:Q], 9)
        raw = cport.box_raw(_f32(vs).astype(np.float32), _f32(ret).astype(np.float32), DIST, np.nan_to_num(q[:Q], nan=1e30), 9)
        assert np.array_equal(mom[:, :, 0], want[:, :, 0]), Q
        assert np.array_equal(mom[:, :, 0], raw[:, :, 0]) and np.abs(mom[:, :, 1:] - raw[:, :, 1:]).max() < 1e-9
    assert mom[:60, :, 0].sum() > 500 and mom[62:65].sum() == 0


def test_many_queries_on_a_buffer_of_many_tiles_per_cta():
    """The tile ring of the filtered scan wraps several times per CTA (700 001 rows = 2 735 tiles over at most
    296 CTAs, ragged last tile), with the bins in global memory (64 x 40 > 1024) and in shared memory (64 x 8):
    counts equal the C oracle's on the same float32 rows; the fused launch gives the same totals."""
    rng = np.random.default_rng(5)
    M, na = 700_001, 40
    centers = rng.normal(0, 1.5, (3000, 20)) * DIST[:20]
    which = rng.integers(0, len(centers), M)
    vs = np.hstack([centers[which] + rng.uniform(-0.6, 0.6, (M, 20)) * DIST[:20],
                    rng.integers(0, na, (M, 1)).astype(np.float64)]).astype(np.float32)
    ret = (-0.95 ** rng.integers(0, 10, M) * (rng.random(M) < 0.5)).astype(np.float32)
    buf = dc.BoxBuffer(20, capacity=M)
    buf.append(vs, ret)
    q = centers[rng.integers(0, len(centers), 64)] + rng.uniform(-0.3, 0.3, (64, 20)) * DIST[:20]
    from dcarl_b200 import dist as dd
    for n_actions in (na, 8):
        raw = cport.box_raw(vs, ret, DIST, q, n_actions, threads=8)
        mom = buf.moments(q, n_actions).cpu().numpy()
        assert np.array_equal(mom[:, :, 0], raw[:, :, 0]) and raw[:, :, 0].sum() > 2000
        assert np.abs(mom[:, :, 1:] - raw[:, :, 1:]).max() < 1e-9
        fq = dd.ShardedBoxQuery(buf, 64, n_actions)
        for rep in range(3):                                               # the two accumulators alternate between calls
            fm, _, _ = fq(q)
            assert np.array_equal(fm[:, :, 0].cpu().numpy(), raw[:, :, 0]), rep
        fm, _, _ = fq(q[::-1].copy())
        assert np.array_equal(fm[:, :, 0].cpu().numpy(), raw[::-1, :, 0])
        one = buf.moments(q[:1], n_actions).cpu().numpy()                  # the plain scan agrees
        assert np.array_equal(one[0, :, 0], raw[0, :, 0])


def test_batched_add_data_matches_the_reference_deque():
    """SURVEY next-row f2: many transition streams at once against the literal deque restatement
    (RLS.py:185-215), including streams with no done, several dones, and fewer than 10 steps."""
    rng = np.random.default_rng(2)
    lens = [0, 1, 9, 10, 11, 25, 40, 200] + rng.integers(1, 120, 40).tolist()
    obs, act, rew, done, offs = [], [], [], [], [0]
    want_state, want_ret = [], []
    for e, T in enumerate(lens):
        o = rng.normal(0, 5, (T, 20)).astype(np.float32)
        a = rng.integers(0, 8, T).astype(np.int32)
        r = np.where(rng.random(T) < 0.2, -rng.random(T), 0.0).astype(np.float32)
        d = rng.random(T) < (0.0 if e % 5 == 0 else 0.03)
        if T and e % 3 == 1:
            d[-1] = True
        rec = pyref.add_data_returns(r.astype(np.float64).tolist(), d.tolist())
        assert [t for t, _ in rec] == list(range(len(rec)))              # recorded rows are a prefix, in time order
        for t, R in rec:
            want_state.append(np.append(o[t], np.float32(a[t])))
            want_ret.append(np.float32(R))
        obs.append(o); act.append(a); rew.append(r); done.append(d); offs.append(offs[-1] + T)
    buf = dc.BoxBuffer(20, capacity=8)
    n = buf.append_episodes(np.concatenate(obs), np.concatenate(act), np.concatenate(rew), np.concatenate(done), offs)
    assert n == len(want_ret) == len(buf)
    assert np.array_equal(buf.state.cpu().numpy(), np.asarray(want_state, np.float32))
    assert np.allclose(buf.ret.cpu().numpy(), np.asarray(want_ret, np.float32), rtol=2e-7, atol=0)
    # appending again lands after the first block
    m = buf.append_episodes(obs[7], act[7], rew[7], done[7], [0, lens[7]])
    assert len(buf) == n + m and m >= lens[7] - 10


def test_sorted_index_gives_the_same_answers_on_a_fraction_of_the_rows(golden):
    """SURVEY next-row f1: the on-device index only narrows the scanned row range."""
    g = golden("rls_box")
    vs, vv, dr = g["visited_state"], g["visited_value"], g["driving_record"]
    plain = dc.BoxBuffer(20)
    plain.append(vs, vv[:, 1])
    idx = dc.BoxBuffer(20)
    idx.append(vs[:-5001], vv[:-5001, 1])
    dim = idx.build_index()
    assert 0 <= dim < 20
    idx.append(vs[-5001:], vv[-5001:, 1])                      # unsorted tail, odd length
    for lo in range(0, 107, 8):
        q = dr[lo:lo + 8, :20]
        n0, m0, v0, p0 = plain.query(q, 8)
        n1, m1, v1, p1 = idx.query(q, 8)
        assert np.array_equal(n0, n1) and np.array_equal(p0, p1)
        assert np.abs(m0 - m1).max() < 1e-12 and np.abs(v0 - v1).max() < 1e-12
        assert idx.last_scanned_rows < len(q) * len(vs)       # the fixture keeps only near-miss rows: low selectivity
    assert np.array_equal(idx.query(dr[:, :20], 8)[0][:, 0], dr[:, 23])     # > 8 queries: the single full pass
    # all 107 logged queries from their index ranges in ONE launch (+ the unsorted tail), more candidate actions too
    for na in (8, 40):
        mi = idx.moments_indexed(dr[:, :20], na).cpu().numpy()
        mp = plain.moments(dr[:, :20], na).cpu().numpy()
        assert np.array_equal(mi[:, :, 0], mp[:, :, 0]) and np.abs(mi - mp).max() < 1e-9
    assert np.array_equal(mi[:, 0, 0], dr[:, 23]) and idx.last_scanned_rows < 107 * len(vs)
    # spread-out data: a query touches a small fraction of the buffer
    rng = np.random.default_rng(31)
    M = 200_000
    wide = np.hstack([rng.normal(0, 60, (M, 20)) * DIST[:20], rng.integers(0, 8, (M, 1)).astype(np.float64)])
    wret = -rng.random(M)
    wb = dc.BoxBuffer(20, capacity=M)
    wb.append(wide, wret)
    wb.build_index()
    qs = wide[rng.integers(0, M, 4), :20]
    n, *_ = wb.query(qs, 8)
    assert wb.last_scanned_rows < 0.05 * 4 * M
    assert np.array_equal(n, _oracle_nmv(wide, wret, qs, 8)[:, :, 0]) and n.sum() >= 4
    for d in (0, 5, 17):                                       # any dimension is a valid sort key
        b = dc.BoxBuffer(20)
        b.append(vs, vv[:, 1])
        assert b.build_index(d) == d
        assert np.array_equal(b.query(dr[40:44, :20], 8)[0], plain.query(dr[40:44, :20], 8)[0])


def test_file_ingest_round_trip(tmp_path, golden):
    g = golden("rls_box")
    vs, vv, dr = g["visited_state"][:4000], g["visited_value"][:4000], g["driving_record"]
    np.savetxt(tmp_path / "visited_state.txt", vs, fmt="%f")
    np.savetxt(tmp_path / "visited_value.txt", vv, fmt="%f")
    buf = dc.load_rls_buffer(str(tmp_path))
    assert len(buf) == 4000
    n0 = buf.query(dr[:5, :20], 8)[0]
    out = tmp_path / "copy"
    out.mkdir()
    dc.save_rls_buffer(buf, str(out))
    again = dc.load_rls_buffer(str(out), build_index=False)
    assert np.array_equal(again.query(dr[:5, :20], 8)[0], n0)
    s2 = golden("sim2")
    data = np.stack([s2["idx"], s2["state"], s2["act"], s2["R"]], 1).astype(np.float64)
    np.save(tmp_path / "data.npy", data)
    idx, act, R = dc.load_sim_data(str(tmp_path / "data.npy"), n_rows=20000)
    res = dc.stream_bounds(idx, act, R, 20, 11)
    assert np.array_equal(res.act, s2["step_act"])
