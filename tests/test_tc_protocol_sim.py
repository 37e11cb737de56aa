"""Randomised simulation of the mbarrier protocols of the two tcgen05 attention kernels (swin_window_attn_tc.cu,
mha_tc.cu): every role (loader warps, the MMA lane, softmax warps) is a coroutine issuing the same wait / arrive / commit
sequence, with the same counts and phase parities, as the CUDA code; tcgen05.commit arrivals are delivered asynchronously
but in order (the tensor pipe retires in issue order).  A wrong parity, count or ordering shows up as a deadlock or as a
wait that is overtaken by two phase completions.  This checks the protocol, not the hardware."""
import random

import pytest


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.phase, self.completions = count, count, 0, 0

    def arrive(self, n=1):
        for _ in range(n):
            self.pending -= 1
            assert self.pending >= 0
            if self.pending == 0:
                self.phase ^= 1
                self.pending = self.count
                self.completions += 1

    def ready(self, parity):          # mbarrier.try_wait.parity: the phase with this parity has completed
        return self.phase != parity


def simulate(roles, bars, rng, max_steps=2_000_000):
    """roles: dict name -> generator yielding ("wait", bar, parity, expected_completion) | ("arrive", bar, n) |
    ("commit", [bars]).  One random enabled action per step; returns when all roles finish, raises on deadlock."""
    pipe = []                          # in-order queue of pending tcgen05.commit arrivals
    live = dict(roles)
    pending = {}                       # name -> the operation the role is about to execute
    for _ in range(max_steps):
        for name in list(live):
            if name not in pending:
                try:
                    pending[name] = next(live[name])
                except StopIteration:
                    del live[name]
        if not live and not pipe:
            return
        enabled = []
        for name, op in pending.items():
            if op[0] == "wait":
                _, b, parity, want = op
                bar = bars[b]
                # the completion this wait is meant for must not have been overtaken (parity aliasing)
                assert bar.completions <= want + 1, f"{name}: wait on {b} for completion {want} overtaken ({bar.completions})"
                if bar.ready(parity):
                    assert bar.completions == want + 1, f"{name}: {b} parity {parity} passes at completion {bar.completions}, wanted {want + 1}"
                    enabled.append(name)
            else:
                enabled.append(name)
        if pipe:
            enabled.append(None)       # the tensor pipe retires its oldest commit
        if not enabled:
            raise AssertionError(f"deadlock: {pending}")
        name = rng.choice(enabled)
        if name is None:
            for b in pipe.pop(0):
                bars[b].arrive()
            continue
        op = pending.pop(name)
        if op[0] == "arrive":
            bars[op[1]].arrive(op[2])
        elif op[0] == "commit":
            pipe.append(list(op[1]))
    raise AssertionError("simulation did not terminate")


def W(bar, completion):                # wait for the `completion`-th (0-based) phase completion of `bar`
    return ("wait", bar, completion & 1, completion)


# ---- swin_window_attn_tc.cu --------------------------------------------------------------------------------------------
def wintc_roles(count):
    bars = {f"QKV_FULL{s}": Bar(128) for s in range(2)}
    bars.update({f"QKV_EMPTY{s}": Bar(1) for s in range(2)})
    for t, nwarp, nthr in ((0, 8, 256), (1, 1, 32)):
        bars[f"S_FULL{t}"], bars[f"S_FREE{t}"] = Bar(1), Bar(nwarp)
        bars[f"P_FULL{t}"], bars[f"P_FREE{t}"] = Bar(nthr), Bar(1)
        bars[f"O_FULL{t}"], bars[f"O_FREE{t}"] = Bar(1), Bar(nwarp)

    def loader(warp):
        for it in range(count):
            s = it & 1
            if it >= 2:
                yield W(f"QKV_EMPTY{s}", (it >> 1) - 1)
            yield ("arrive", f"QKV_FULL{s}", 32)

    def mma():
        for n in range(-1, count):
            if n + 1 < count:
                m, s = n + 1, (n + 1) & 1
                yield W(f"QKV_FULL{s}", m >> 1)
                for t in range(2):
                    if m > 0:
                        yield W(f"S_FREE{t}", m - 1)
                    yield ("commit", [f"S_FULL{t}"])
            if n < 0:
                continue
            s = n & 1
            for t in range(2):
                yield W(f"P_FULL{t}", n)
                if n > 0:
                    yield W(f"O_FREE{t}", n - 1)
                yield ("commit", [f"O_FULL{t}", f"P_FREE{t}"] + ([f"QKV_EMPTY{s}"] if t == 1 else []))

    def softmax(t):
        def unit(n):
            yield W(f"S_FULL{t}", n)
            yield ("arrive", f"S_FREE{t}", 1)
            if n > 0:
                yield W(f"P_FREE{t}", n - 1)
            yield ("arrive", f"P_FULL{t}", 32)

        def epilogue(n):
            yield W(f"O_FULL{t}", n)
            yield ("arrive", f"O_FREE{t}", 1)

        if count <= 0:
            return
        yield from unit(0)
        for n in range(count):
            if n + 1 < count:
                yield from unit(n + 1)
            yield from epilogue(n)

    roles = {f"loader{w}": loader(w) for w in range(4)}
    roles["mma"] = mma()
    roles.update({f"softmax0_{w}": softmax(0) for w in range(8)})
    roles["softmax1"] = softmax(1)
    return roles, bars


# ---- mha_tc.cu ----------------------------------------------------------------------------------------------------------
def mhatc_roles(nblocks, stages=3):
    bars = {f"KV_FULL{s}": Bar(128) for s in range(stages)}
    bars.update({f"KV_EMPTY{s}": Bar(1) for s in range(stages)})
    bars.update(Q_FULL=Bar(128), S_FULL=Bar(1), S_FREE=Bar(8), P_FULL=Bar(256), P_FREE=Bar(1), O_FULL=Bar(1), O_FREE=Bar(8))
    count = 2 * nblocks

    def loader(warp):
        yield ("arrive", "Q_FULL", 32)
        for it in range(nblocks):
            s, use = it % stages, it // stages
            if use > 0:
                yield W(f"KV_EMPTY{s}", use - 1)
            yield ("arrive", f"KV_FULL{s}", 32)

    def mma():
        yield W("Q_FULL", 0)
        for n in range(-1, count):
            if n + 1 < count:
                m = n + 1
                blk, tile = m >> 1, m & 1
                s = blk % stages
                if tile == 0:
                    yield W(f"KV_FULL{s}", blk // stages)
                if m > 0:
                    yield W("S_FREE", m - 1)
                yield ("commit", ["S_FULL"])
            if n < 0:
                continue
            blk, tile = n >> 1, n & 1
            s = blk % stages
            yield W("P_FULL", n)
            if n > 0:
                yield W("O_FREE", n - 1)
            yield ("commit", ["O_FULL", "P_FREE"] + ([f"KV_EMPTY{s}"] if tile == 1 else []))

    def softmax(warp):
        def unit(n):
            yield W("S_FULL", n)
            yield ("arrive", "S_FREE", 1)
            if n > 0:
                yield W("P_FREE", n - 1)
            yield ("arrive", "P_FULL", 32)

        def merge(n):
            yield W("O_FULL", n)
            yield ("arrive", "O_FREE", 1)

        if nblocks > 0:
            yield from unit(0)
        for i in range(nblocks):
            yield from unit(2 * i + 1)
            yield from merge(2 * i)
            if i + 1 < nblocks:
                yield from unit(2 * i + 2)
            yield from merge(2 * i + 1)

    roles = {f"loader{w}": loader(w) for w in range(4)}
    roles["mma"] = mma()
    roles.update({f"softmax{w}": softmax(w) for w in range(8)})
    return roles, bars


@pytest.mark.parametrize("count", [1, 2, 3, 4, 7])
def test_window_kernel_protocol(count):
    for seed in range(25):
        roles, bars = wintc_roles(count)
        simulate(roles, bars, random.Random(seed))
        assert bars["O_FULL0"].completions == count and bars["O_FULL1"].completions == count


@pytest.mark.parametrize("nblocks", [1, 2, 3, 4, 8])
def test_mha_kernel_protocol(nblocks):
    for seed in range(25):
        roles, bars = mhatc_roles(nblocks)
        simulate(roles, bars, random.Random(seed))
        assert bars["O_FULL"].completions == 2 * nblocks


def test_simulator_detects_a_wrong_parity():
    """sanity of the checker itself: dropping the S_FREE wait lets S(n+1) overtake the softmax"""
    roles, bars = mhatc_roles(3)

    def bad_mma():
        yield W("Q_FULL", 0)
        for m in range(6):
            if m & 1 == 0:
                yield W(f"KV_FULL{(m >> 1) % 3}", (m >> 1) // 3)
            yield ("commit", ["S_FULL"])          # no S_FREE wait
        for n in range(6):
            yield W("P_FULL", n)
            yield ("commit", ["O_FULL", "P_FREE"] + ([f"KV_EMPTY{(n >> 1) % 3}"] if n & 1 else []))

    roles["mma"] = bad_mma()
    with pytest.raises(AssertionError):
        for seed in range(50):
            r, b = mhatc_roles(3)
            r["mma"] = bad_mma()
            simulate(r, b, random.Random(seed))
