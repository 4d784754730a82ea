"""GPU tests of the widened rows: sessions, in-memory chaining of operators for whole
postfix expressions, and batched requests — against the oracle's cloud main() chained the way the reference
chains it, and against Python integers."""
import os

import numpy as np
import pytest

import oracle_bind as ob

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(pkg, oracle):
    eng = pkg.Engine(0)
    ks = oracle.keygen(ob.params_default(8), seed=31)
    nbit = oracle.keygen(ob.params_default(8), seed=32)
    key = eng.cloud_key_from_arrays(pkg.Params.default(8), ks.bk_coef(), ks.ksk())
    sess = eng.session_from_keys(key, nbit.lwe_key())
    yield eng, ks, nbit, sess
    sess.close(); key.close(); ks.free(); nbit.free(); eng.close()


def _signed(sign, v):
    return -v if sign == 2 else v


def test_session_compute_matches_oracle_main(oracle, env):
    eng, ks, nbit, sess = env
    for op, s1, s2, width, a, b in [(1, 0, 0, 32, 1 << 30, 1 << 30), (2, 0, 0, 32, 7, 1000), (4, 2, 0, 32, 77777, 99999),
                                    (1, 2, 2, 64, 1 << 40, 12345)]:
        A, B = oracle.alice(ks, nbit, s1, width, a, seed=1), oracle.alice(ks, nbit, s2, width, b, seed=2)
        rc, ans, _ = sess.compute(op, A, B)
        rc_o, ans_o = oracle.cloud_main(ks, nbit, op, np.concatenate([A, B]))
        assert rc == rc_o == 0 and len(ans) == 352
        assert oracle.verif(ks, nbit, ans) == oracle.verif(ks, nbit, ans_o)
        code, w, chunks = oracle.verif(ks, nbit, ans)
        assert ob.decode_result(op, code, w, chunks) == {1: lambda x, y: x + y, 2: lambda x, y: x - y, 4: lambda x, y: x * y}[op](
            _signed(s1, a), _signed(s2, b))


@pytest.mark.parametrize("postfix,fn", [("AB*C+", lambda a, b, c: a * b + c), ("ABC*+", lambda a, b, c: a + b * c),
                                         ("AB+C+", lambda a, b, c: a + b + c), ("AB+C-", lambda a, b, c: a + b - c)])
def test_postfix_chaining(oracle, env, postfix, fn):
    """3-operand expressions (BASELINE.json config 3; the paper's A+B+C, A+B-C, A+B*C): the GPU session against
    the oracle chained through memory the way dragonfly_cipher_cloud.py chains through files."""
    eng, ks, nbit, sess = env
    vals = [(123456, 7890, 4242), (1 << 30, 1 << 30, 1 << 30)]
    ops = np.stack([np.stack([oracle.alice(ks, nbit, 0, 32, v, seed=10 * e + k) for k, v in enumerate(t)]) for e, t in enumerate(vals)])
    rc, ans, counts, secs = sess.eval_postfix(postfix, ops)
    assert rc == 0 and (counts == 352).all()
    for e, t in enumerate(vals):
        # oracle: same walk, operator by operator
        stack = []
        for tok in postfix:
            if tok.isalpha():
                stack.append(ops[e, ord(tok) - 65])
            else:
                y, x = stack.pop(), stack.pop()
                _, r = oracle.cloud_main(ks, nbit, {"+": 1, "-": 2, "*": 4}[tok], np.concatenate([x, y]))
                stack.append(r)
        want = oracle.verif(ks, nbit, stack[-1])
        got = oracle.verif(ks, nbit, ans[e])
        assert got == want
        code, w, chunks = got
        last_op = {"+": 1, "-": 2, "*": 4}[postfix[-1]]
        assert ob.decode_result(last_op, code, w, chunks) == fn(*t)


def test_batched_requests_mixed_circuits(oracle, env):
    """f-4: a batch of independent requests with different operators, signs and widths in one call"""
    eng, ks, nbit, sess = env
    reqs = [(1, 0, 0, 32, 11, 22), (2, 0, 0, 32, 100, 1), (1, 0, 0, 32, 5, 6), (4, 0, 0, 32, 300, 400), (2, 0, 2, 32, 9, 3),
            (1, 0, 0, 64, 1 << 35, 7), (2, 0, 0, 32, 3, 100), (4, 0, 0, 256, 1, 1)]
    o1 = np.stack([oracle.alice(ks, nbit, r[1], r[3], r[4], seed=2 * i) for i, r in enumerate(reqs)])
    o2 = np.stack([oracle.alice(ks, nbit, r[2], r[3], r[5], seed=2 * i + 1) for i, r in enumerate(reqs)])
    codes, ans, counts, _ = sess.compute_batch([r[0] for r in reqs], o1, o2)
    for i, (op, s1, s2, width, a, b) in enumerate(reqs):
        if op == 4 and width == 256:
            assert codes[i] == 126 and counts[i] == 64          # Cloud/cloud.c:860-864
            continue
        assert codes[i] == 0 and counts[i] == 352
        code, w, chunks = oracle.verif(ks, nbit, ans[i])
        want = {1: _signed(s1, a) + _signed(s2, b), 2: _signed(s1, a) - _signed(s2, b), 4: _signed(s1, a) * _signed(s2, b)}[op]
        assert ob.decode_result(op, code, w, chunks) == want, i


def test_cloud_dynamic_mirror(tmp_path, pkg, oracle, env):
    """the ctypes replacement for dragonfly_cipher_cloud.py's compute()/compute_final() on the reference's files"""
    eng, ks, nbit, _ = env
    from ieache_b200 import cloud_dynamic as cd
    d = str(tmp_path)
    ks.write_cloud_key(os.path.join(d, "cloud.key"))
    nbit.write_secret_key(os.path.join(d, "nbit.key"))
    files = []
    for k, v in enumerate((1000, 2000, 3000)):
        p = os.path.join(d, f"client{k}.data")
        ks.write_samples(oracle.alice(ks, nbit, 0, 32, v, seed=k), p)
        files.append(p)
    node = cd.CloudNode(d, engine=eng)
    # compute(): operator on cloud.data = client0 || client1
    blk = np.concatenate([cd.read_block(files[0], 8), cd.read_block(files[1], 8)])
    cd.write_block(os.path.join(d, "cloud.data"), blk, node.variance)
    assert os.path.getsize(os.path.join(d, "cloud.data")) == 704 * (4 + 4 * 9 + 8)
    assert node.compute(1) == 0
    code, w, chunks = oracle.verif(ks, nbit, ks.read_samples(os.path.join(d, "answer.data"), 352))
    assert (code, w, chunks[0]) == (0, 32, 3000)
    # evaluate(): A*B+C
    assert node.evaluate("AB*C+", files) == 0
    code, w, chunks = oracle.verif(ks, nbit, ks.read_samples(os.path.join(d, "answer.data"), 352))
    assert w == 64 and chunks[0] | (chunks[1] << 32) == 1000 * 2000 + 3000
    node.close()


def test_chained_operators_keep_values_on_the_device(oracle, env):
    """in a postfix expression every operand goes to the GPU once and only the final answers come back;
    intermediate results never visit the host (the context counts the 352-sample blocks the session calls copy).
    Five operators over four operands, 70 instances (two passes of 64): 4 x 70 uploads, 70 downloads."""
    eng, ks, nbit, sess = env
    n_expr = 70
    rng = np.random.default_rng(70)
    vals = rng.integers(1, 1 << 12, size=(n_expr, 4))
    ops = np.stack([np.stack([oracle.alice(ks, nbit, 0, 32, int(v), seed=100 * e + k) for k, v in enumerate(t)]) for e, t in enumerate(vals)])
    h0, d0 = eng.copy_counts()
    rc, ans, counts, _ = sess.eval_postfix("AB+C+D+A-B+", ops)
    h1, d1 = eng.copy_counts()
    assert rc == 0 and (counts == 352).all()
    assert (h1 - h0, d1 - d0) == (4 * n_expr, n_expr)
    for e, (a, b, c, d) in enumerate(vals):
        code, w, chunks = oracle.verif(ks, nbit, ans[e])
        assert ob.decode_result(1, code, w, chunks) == int(a + b + c + d - a + b), e


def test_sign_code_of_a_chained_product_of_negatives(oracle, env):
    """(-a) * (-b) carries sign code 4; added to c the reference writes code 0 for the sum 4 + 0 (cloud.c:812-821) and runs
    the add branch: the chained session must do what the reference's cloud.c does with that operand block"""
    eng, ks, nbit, sess = env
    a, b, c = 1234, 77, 99
    ops = np.stack([oracle.alice(ks, nbit, 2, 32, a, seed=1), oracle.alice(ks, nbit, 2, 32, b, seed=2), oracle.alice(ks, nbit, 0, 32, c, seed=3)])[None]
    rc, ans, counts, _ = sess.eval_postfix("AB*C+", ops)
    assert rc == 0
    _, prod = oracle.cloud_main(ks, nbit, 4, np.concatenate([ops[0, 0], ops[0, 1]]))
    assert oracle.verif(ks, nbit, prod)[0] == 4
    _, want = oracle.cloud_main(ks, nbit, 1, np.concatenate([prod, ops[0, 2]]))
    assert oracle.verif(ks, nbit, ans[0]) == oracle.verif(ks, nbit, want)
    assert oracle.verif(ks, nbit, ans[0])[0] == 0


def test_ingest_of_many_request_directories_in_bounded_passes(tmp_path, pkg, oracle, env):
    """SURVEY 8 f-4 at a size that needs several passes: 150 request directories (3 passes of 64) through
    compute_dirs; host and pinned memory are bounded by the pass size, not by the number of directories"""
    eng, ks, nbit, _ = env
    keys = str(tmp_path / "keys")
    os.makedirs(keys)
    ks.write_cloud_key(os.path.join(keys, "cloud.key"))
    nbit.write_secret_key(os.path.join(keys, "nbit.key"))
    sess = eng.session(os.path.join(keys, "cloud.key"), os.path.join(keys, "nbit.key"))
    dirs, want = [], []
    for k in range(150):
        d = str(tmp_path / f"r{k}")
        os.makedirs(d)
        op = (1, 2, 1)[k % 3]
        a, b = 1000 + 7 * k, 3 * k + 1
        data = np.concatenate([oracle.alice(ks, nbit, 0, 32, a, seed=2 * k), oracle.alice(ks, nbit, 0, 32, b, seed=2 * k + 1)])
        ks.write_samples(data, os.path.join(d, "cloud.data"))
        open(os.path.join(d, "operator.txt"), "w").write(str(op))
        dirs.append(d); want.append(a + b if op == 1 else a - b)
    assert sess.set_pass(64) == 256               # default pass size; 64 makes this three passes with read-ahead / write-behind
    codes, secs = sess.compute_dirs(dirs)
    assert (codes == 0).all() and secs > 0
    for d, k in zip(dirs, range(150)):
        code, w, chunks = oracle.verif(ks, nbit, ks.read_samples(os.path.join(d, "answer.data"), 352))
        assert ob.decode_result((1, 2, 1)[k % 3], code, w, chunks) == want[k], k
    sess.close()
