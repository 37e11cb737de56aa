"""Randomised simulation of the cluster protocol of csrc/mask_einsum_mc.cu (two CTAs, multicast TMA, multicast
tcgen05.commit) with DATA-HAZARD tracking: every shared-memory stage half, the resident E tile and the TMEM accumulators
carry a version; asynchronous agents (each CTA's TMA engine: unordered; each CTA's tensor pipe: in order) deliver their
effects at random later times.  Checked: no deadlock, no wait overtaken by two phase completions, no operand overwritten
while an MMA that reads it is still in flight (WAR), every MMA reads the versions it was issued for (RAW), every epilogue
reads the accumulator of its own tile.  Also: the query split and the tile ranges of the launcher.  The kernel itself has
never run on hardware; this checks the protocol, not the hardware."""
import random

import pytest

STAGES = 6


class Bar:
    """mbarrier with arrival count and transaction bytes; a phase completes when both are drained"""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase, self.completions = count, count, 0, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count
            self.completions += 1

    def arrive(self, expect_tx=0):
        assert self.pending > 0, "more arrivals than the barrier's count in one phase"
        self.tx += expect_tx
        self.pending -= 1
        self._check()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._check()

    def ready(self, parity):
        return self.phase != parity


def query_split(Q, rank):
    """einmc::query_split"""
    npad = (Q + 15) & ~15
    first = ((npad // 16 + 1) // 2) * 16
    n0 = 0 if rank == 0 else first
    n = first if rank == 0 else npad - first
    valid = max(min(n0 + n, Q) - n0, 0)
    return n0, n, valid


def tile_range(num_tiles, clusters, c):
    return num_tiles * c // clusters, num_tiles * (c + 1) // clusters


class Cluster:
    """state of one CTA pair + the role coroutines, following the kernel statement by statement"""

    def __init__(self, tiles, tiles_per_frame, kchunks, rng):
        self.tiles, self.tpf, self.kchunks, self.rng = tiles, tiles_per_frame, kchunks, rng
        self.bars = {}
        for r in (0, 1):
            for s in range(STAGES):
                self.bars[(r, "FULL", s)] = Bar(1)
                self.bars[(r, "EMPTY", s)] = Bar(2)
            self.bars[(r, "E_FULL", 0)] = Bar(1)
            self.bars[(r, "E_EMPTY", 0)] = Bar(1)
            for a in (0, 1):
                self.bars[(r, "T_FULL", a)] = Bar(1)
                self.bars[(r, "T_EMPTY", a)] = Bar(8)
        # data versions: F stage halves per CTA (what (tile, kc) they hold), resident E (frame), TMEM accumulators (tile)
        self.f = {(r, s, h): None for r in (0, 1) for s in range(STAGES) for h in (0, 1)}
        self.e = {0: None, 1: None}
        self.acc = {(r, a): None for r in (0, 1) for a in (0, 1)}
        self.tma = []                                  # unordered pending TMA deliveries
        self.pipe = {0: [], 1: []}                     # in-order tensor pipes: ("mma", reads, acc) | ("commit", bars)
        self.stores = {0: [], 1: []}                   # tiles written out by each CTA (by warp 4 of its epilogue)

    # ---- asynchronous effects --------------------------------------------------------------------------------------
    def _reads_in_flight(self, r):
        return [item for item in self.pipe[r] if item[0] == "mma"]

    def deliver_tma(self, item):
        kind = item[0]
        if kind == "F":                                # multicast: one half-stage into both CTAs
            _, src, s, h, version, nbytes = item
            for r in (0, 1):
                for m in self._reads_in_flight(r):
                    assert (s, h) not in [(x[0], x[1]) for x in m[1]["f"]], \
                        f"WAR: CTA{src} refills stage {s} half {h} while an MMA of CTA{r} still reads it"
                self.f[(r, s, h)] = version
                self.bars[(r, "FULL", s)].complete_tx(nbytes)
        else:                                          # "E": this CTA's resident half of E
            _, r, frame, nbytes = item
            for m in self._reads_in_flight(r):
                assert m[1]["e"] is None or m[1]["e"] == frame, f"WAR: CTA{r} replaces E while an MMA still reads frame {m[1]['e']}"
            self.e[r] = frame
            self.bars[(r, "E_FULL", 0)].complete_tx(nbytes)

    def retire(self, r):
        item = self.pipe[r].pop(0)
        if item[0] == "mma":
            _, reads, acc, tile = item
            for (s, h, version) in reads["f"]:
                assert self.f[(r, s, h)] == version, f"RAW: CTA{r} MMA of {version} found {self.f[(r, s, h)]} in stage {s} half {h}"
            assert self.e[r] == reads["e"], f"RAW: CTA{r} MMA of frame {reads['e']} found E of frame {self.e[r]}"
            self.acc[(r, acc)] = tile
        else:
            for key in item[1]:
                self.bars[key].arrive()

    # ---- roles ---------------------------------------------------------------------------------------------------------
    def producer(self, r):
        stage, use, frame_loaded, e_loads = 0, 0, -1, 0
        for tile in self.tiles:
            t = tile // self.tpf
            if t != frame_loaded:
                if e_loads > 0:
                    yield ("wait", (r, "E_EMPTY", 0), e_loads - 1)
                yield ("expect", (r, "E_FULL", 0), 2 * self.kchunks * 100)
                for _ in range(2 * self.kchunks):
                    yield ("tma", ("E", r, t, 100))
                frame_loaded, e_loads = t, e_loads + 1
            for kc in range(self.kchunks):
                if use > 0:
                    yield ("wait", (r, "EMPTY", stage), use - 1)      # parity trick: the first pass never blocks
                yield ("expect", (r, "FULL", stage), 2 * 8192)
                yield ("tma", ("F", r, stage, r, (tile, kc), 8192))    # CTA r fetches half r of the stage, for both CTAs
                stage += 1
                if stage == STAGES:
                    stage, use = 0, use + 1

    def mma(self, r):
        stage, use, acc, acc_use, frame_ready, e_uses = 0, 0, 0, 0, -1, 0
        for i, tile in enumerate(self.tiles):
            t = tile // self.tpf
            if t != frame_ready:
                yield ("wait", (r, "E_FULL", 0), e_uses)
                frame_ready, e_uses = t, e_uses + 1
            last_of_frame = i + 1 == len(self.tiles) or self.tiles[i + 1] // self.tpf != t
            if acc_use > 0:
                yield ("wait", (r, "T_EMPTY", acc), acc_use - 1)
            for kc in range(self.kchunks):
                yield ("wait", (r, "FULL", stage), use)
                reads = {"f": [(stage, 0, (tile, kc)), (stage, 1, (tile, kc))], "e": t}
                yield ("mma", r, reads, acc, tile)
                yield ("commit", r, [(0, "EMPTY", stage), (1, "EMPTY", stage)])          # multicast commit
                if kc == self.kchunks - 1:
                    yield ("commit", r, [(r, "T_FULL", acc)])
                    if last_of_frame:
                        yield ("commit", r, [(r, "E_EMPTY", 0)])
                stage += 1
                if stage == STAGES:
                    stage, use = 0, use + 1
            acc += 1
            if acc == 2:
                acc, acc_use = 0, acc_use + 1

    def epilogue(self, r, w):
        acc, acc_use = 0, 0
        for tile in self.tiles:
            yield ("wait", (r, "T_FULL", acc), acc_use)
            yield ("read_acc", r, acc, tile, w)
            yield ("arrive", (r, "T_EMPTY", acc))
            acc += 1
            if acc == 2:
                acc, acc_use = 0, acc_use + 1

    def run(self, max_steps=3_000_000):
        roles = {}
        for r in (0, 1):
            roles[f"prod{r}"] = self.producer(r)
            roles[f"mma{r}"] = self.mma(r)
            for w in range(8):
                roles[f"epi{r}_{w}"] = self.epilogue(r, w)
        pending = {}
        for _ in range(max_steps):
            for name in list(roles):
                if name not in pending:
                    try:
                        pending[name] = next(roles[name])
                    except StopIteration:
                        del roles[name]
            if not roles and not self.tma and not self.pipe[0] and not self.pipe[1]:
                return
            enabled = []
            for name, op in pending.items():
                if op[0] == "wait":
                    bar = self.bars[op[1]]
                    want = op[2]
                    assert bar.completions <= want + 1, f"{name}: wait on {op[1]} for completion {want} overtaken ({bar.completions})"
                    if bar.ready(want & 1):
                        assert bar.completions == want + 1
                        enabled.append(name)
                else:
                    enabled.append(name)
            enabled += [("tma", i) for i in range(len(self.tma))]
            enabled += [("pipe", r) for r in (0, 1) if self.pipe[r]]
            if not enabled:
                raise AssertionError(f"deadlock: {pending}")
            pick = self.rng.choice(enabled)
            if isinstance(pick, tuple):
                if pick[0] == "tma":
                    self.deliver_tma(self.tma.pop(pick[1]))
                else:
                    self.retire(pick[1])
                continue
            op = pending.pop(pick)
            if op[0] == "expect":
                self.bars[op[1]].arrive(expect_tx=op[2])
            elif op[0] == "arrive":
                self.bars[op[1]].arrive()
            elif op[0] == "tma":
                self.tma.append(op[1])
            elif op[0] == "mma":
                self.pipe[op[1]].append(("mma", op[2], op[3], op[4]))
            elif op[0] == "commit":
                self.pipe[op[1]].append(("commit", op[2]))
            elif op[0] == "read_acc":
                _, r, acc, tile, w = op
                assert self.acc[(r, acc)] == tile, f"epilogue of CTA{r} reads accumulator {acc} holding {self.acc[(r, acc)]}, wanted {tile}"
                if w == 0:
                    self.stores[r].append(tile)
        raise AssertionError("simulation did not terminate")


@pytest.mark.parametrize("tiles,tpf,kchunks", [
    (list(range(3)), 10, 2),            # inside one frame, fewer chunk loads than stages
    (list(range(7, 16)), 10, 8),        # crosses one frame boundary (north-star: 8 chunks)
    (list(range(8, 31)), 10, 3),        # crosses two boundaries; 3 chunks per tile walks the 6-stage ring out of phase
    (list(range(0, 1)), 1, 8),          # a single tile
    (list(range(0, 5)), 1, 1),          # every tile a new frame
])
def test_cluster_protocol_and_hazards(tiles, tpf, kchunks):
    for seed in range(12):
        c = Cluster(tiles, tpf, kchunks, random.Random(seed))
        c.run()
        assert c.stores[0] == tiles and c.stores[1] == tiles


def test_simulator_catches_a_single_cta_release():
    """sanity of the checker: if the MMA warp released a stage in its own CTA only, the peer's refill would land under a
    running MMA (or the peer's producer would starve) -- the simulation must fail"""
    class Broken(Cluster):
        def mma(self, r):
            for op in super().mma(r):
                if op[0] == "commit" and len(op[2]) == 2:
                    yield ("commit", r, [(r, "EMPTY", op[2][0][2])] * 2)      # both arrivals on the own barrier
                else:
                    yield op

    failures = 0
    for seed in range(10):
        try:
            Broken(list(range(12)), 10, 8, random.Random(seed)).run()
        except AssertionError:
            failures += 1
    assert failures == 10


def test_query_split_and_tile_ranges():
    for Q in range(17, 257):
        (a0, an, av), (b0, bn, bv) = query_split(Q, 0), query_split(Q, 1)
        assert a0 == 0 and b0 == an and an % 16 == 0 and bn % 16 == 0 and 16 <= bn <= an <= 128
        assert an + bn == (Q + 15) // 16 * 16 and av + bv == Q and av == min(an, Q) and bv >= 1
    assert query_split(200, 0) == (0, 112, 112) and query_split(200, 1) == (112, 96, 88)
    assert query_split(16, 1)[1] == 0                       # the launcher refuses this (falls back to the 1-CTA kernel)
    for tiles, clusters in ((2300, 74), (5, 5), (75, 74), (460, 74)):
        covered = []
        for c in range(clusters):
            b, e = tile_range(tiles, clusters, c)
            assert e - b in (tiles // clusters, tiles // clusters + 1) and e > b
            covered += list(range(b, e))
        assert covered == list(range(tiles))
    # at the north-star shape a cluster sees at most two frames
    for c in range(74):
        b, e = tile_range(2300, 74, c)
        assert (e - 1) // 460 - b // 460 <= 1
