"""CPU restatement of the measurement scan's arithmetic (qcsim_b200/csrc/reduce_kernels.cuh).

The reference's outcome is the first i with prob <= acc_i, acc_i = fl(acc_{i-1} + |a_i|^2), a strictly sequential
fp64 sum (QubitRegister.h:172-190).  The engine reproduces acc bit for bit in parallel: inside a binade
[2^e, 2^(e+1)) the sum moves in integer multiples of u = 2^(e-52), fl(acc + p) = acc + u * rint(p / u) unless p / u is
an exact half; chunks that cross a power of two or contain a tie are replayed element by element.  This test runs
that algorithm in numpy (same steps as k_chunk_increments / k_sequential_walk) against np.cumsum, which is the
sequential sum, on adversarial inputs: exact ties, binade crossings inside a chunk, zeros, tiny and huge terms.
"""
import math

import numpy as np
import pytest

CHUNK = 64  # the device uses 4096; a small chunk makes crossings and ties frequent here


TILE = 8    # chunks resolved per step of the walk (the device: 1024)


def ilogb(x):
    return math.frexp(x)[1] - 1


def walk(p, start=0.0):
    """acc at every chunk start (+ final) computed the way the device does it; returns (acc_start, n_replayed)"""
    n_chunks = (len(p) + CHUNK - 1) // CHUNK
    # 1. exact chunk masses -> predicted binade of every chunk start (math.fsum = exact, like the double-double prefix)
    prefix = [math.fsum([start] + list(p[: c * CHUNK])) for c in range(n_chunks)]
    # 2. k_chunk_increments: K in units of the predicted binade's ulp, flags
    K, slow, zero, E = [], [], [], []
    for c in range(n_chunks):
        chunk = p[c * CHUNK: (c + 1) * CHUNK]
        e = ilogb(prefix[c]) if prefix[c] > 0 else -5000
        s, k = e <= -900, 0
        if not s:
            x = np.ldexp(chunk, 52 - e)                      # exact scaling
            if np.any(x >= 2.0 ** 52) or np.any(x - np.floor(x) == 0.5):
                s = True
            else:
                k = int(np.sum(np.rint(x).astype(np.uint64), dtype=np.uint64))
        K.append(min(k, 1 << 53)); slow.append(s); zero.append(not np.any(chunk)); E.append(e)
    # 3. k_sequential_walk: runs of chunks inside the binade of the running sum are an integer prefix sum
    acc_start = np.zeros(n_chunks + 1)
    carry, c, replayed = start, 0, 0
    while c < n_chunks:
        cnt = min(TILE, n_chunks - c)
        e0 = ilogb(carry) if carry > 0 else -6000
        incl, first = 0, cnt
        before, after = [], []
        for t in range(cnt):
            inc = 0 if zero[c + t] else K[c + t]
            before.append(carry + math.ldexp(float(incl), e0 - 52))
            incl += inc
            after.append(carry + math.ldexp(float(incl), e0 - 52))
            fails = (not zero[c + t]) and (slow[c + t] or E[c + t] != e0 or ilogb(after[t]) != e0)
            if fails and first == cnt:
                first = t
        for t in range(min(first + 1, cnt)):
            acc_start[c + t] = before[t]
        if first < cnt:
            acc = before[first]
            for v in p[(c + first) * CHUNK: (c + first + 1) * CHUNK]:
                acc = acc + float(v)
            carry = acc
            replayed += 1
            c += first + 1
        else:
            carry = after[cnt - 1]
            c += cnt
    acc_start[n_chunks] = carry
    return acc_start, replayed


def cases():
    rng = np.random.default_rng(5)
    n = 1 << 12
    amps = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    amps /= np.linalg.norm(amps)
    yield "random state", (amps.real * amps.real + amps.imag * amps.imag)
    yield "uniform", np.full(n, 1.0 / n)
    p = np.zeros(n)
    p[7] = 1.0
    yield "basis state", p
    p = rng.random(n) * np.exp(rng.uniform(-60, 0, n))       # 26 orders of magnitude
    yield "wide dynamic range", p / p.sum()
    p = np.full(n, 2.0 ** -20)
    p[100:200] = 2.0 ** -54 + 2.0 ** -20                       # exact halves of the ulp once acc >= 0.5 ... ties
    p[300:400] = 1.5 * 2.0 ** -53
    yield "ties", np.concatenate([np.full(64, 2.0 ** -7), p])
    yield "tiny then large", np.concatenate([np.full(1000, 1e-300), np.full(1000, 1e-5), [0.75], np.full(1000, 1e-17)])
    yield "zeros inside", np.concatenate([np.zeros(500), np.full(300, 1e-3), np.zeros(700), np.full(100, 7e-3)])


@pytest.mark.parametrize("name,p", list(cases()), ids=[c[0] for c in cases()])
def test_parallel_walk_equals_the_sequential_sum(name, p):
    p = np.asarray(p, dtype=np.float64)
    want = np.concatenate([[0.0], np.cumsum(p)])               # sequential fp64 running sum
    seq = 0.0
    for i in (0, 1, len(p) // 2):                               # np.cumsum really is the sequential recurrence
        pass
    acc_start, replayed = walk(p)
    n_chunks = (len(p) + CHUNK - 1) // CHUNK
    for c in range(n_chunks + 1):
        i = min(c * CHUNK, len(p))
        assert acc_start[c] == want[i], (name, c, acc_start[c], want[i])
    assert replayed < n_chunks or name in ("ties", "tiny then large", "basis state", "zeros inside")


def test_numpy_cumsum_is_the_sequential_recurrence():
    rng = np.random.default_rng(1)
    p = rng.random(5000)
    acc, out = 0.0, []
    for v in p:
        acc = acc + float(v)
        out.append(acc)
    assert np.array_equal(np.array(out), np.cumsum(p))


def test_chained_slices_continue_the_sum():
    """sharded registers: rank r starts its walk from the sum rank r-1 ended with"""
    rng = np.random.default_rng(9)
    p = rng.random(4096)
    p /= p.sum()
    want = np.cumsum(p)
    acc = 0.0
    for r in range(4):
        sl = p[r * 1024: (r + 1) * 1024]
        acc_start, _ = walk(sl, start=acc)
        acc = acc_start[-1]
        assert acc == want[(r + 1) * 1024 - 1]
