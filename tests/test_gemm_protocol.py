"""Host-side model of the persistent tcgen05 contraction's synchronisation (pcrcg_b200/csrc/gemm_tc.cu, k_gemm_bf16x3): one TMA
producer thread, one MMA thread, EW epilogue warps, a STAGES-deep shared-memory ring and two TMEM accumulator sets, all on
phase-parity mbarriers.  A small configuration is explored EXHAUSTIVELY (every interleaving; TMA completions and tcgen05.commit
arrivals are separate events that may be delayed arbitrarily, commits completing in issue order) and must never

  * let the producer overwrite a stage whose MMAs have not completed,
  * let the MMA thread read a stage before its bytes have landed, or the bytes of another k-block,
  * let the MMAs of tile t + 2 start before every epilogue warp has drained tile t from the same accumulator set,
  * let an epilogue warp read an accumulator that is not the finished tile it expects,
  * deadlock.

The same explorer rejects two deliberately broken variants (a producer that does not wait for the stage to be free; an MMA thread
that does not wait for the accumulator to be drained), so the checks have teeth."""
import re
import os

import pytest


def explore(STAGES, NKB, NT, EW, producer_waits=True, mma_waits_acc=True, limit=3_000_000):
    def test(completed, parity):
        return (completed & 1) != parity

    total = NT * NKB
    # state = (p_it, tma (frozenset of iterations in flight), stage (tuple: None | ("landed", it) | ("loading", it)), consumed (tuple per
    #          stage: last iteration whose MMAs completed, -1), full (tuple of completed counts), empty (tuple), m_it, m_stage,
    #          commits (tuple of pending commits), accfull (tuple 2), accempty (tuple 2 of (completed, pending)),
    #          acc (tuple 2: tile whose result is complete, -1), drained (tuple 2: per acc set, tuple of EW last drained tiles), epi (tuple EW of next tile))
    init = (0, frozenset(), (None,) * STAGES, (-1,) * STAGES, (0,) * STAGES, (0,) * STAGES, 0, 0, (), (0, 0), ((0, EW), (0, EW)), (-1, -1),
            (tuple([-1] * EW),) * 2, (0,) * EW)
    seen, stack, n = set(), [init], 0
    while stack:
        s = stack.pop()
        if s in seen:
            continue
        seen.add(s)
        n += 1
        if n > limit:
            return "limit", n
        (p_it, tma, stage, consumed, full, empty, m_it, m_stage, commits, accfull, accempty, acc, drained, epi) = s
        succ = []

        def rep(t, i, v):
            return t[:i] + (v,) + t[i + 1:]
        # ---- TMA producer: mbar_wait(empty_bar(s), ph ^ 1); expect_tx; 4 tensor loads
        if p_it < total:
            sl, ph = p_it % STAGES, (p_it // STAGES) & 1
            if not producer_waits or test(empty[sl], ph ^ 1):
                if p_it >= STAGES and consumed[sl] != p_it - STAGES:
                    return "violation: producer overwrites a stage in use", (p_it, sl, consumed[sl])
                succ.append((p_it + 1, tma | {p_it}, rep(stage, sl, ("loading", p_it)), consumed, full, empty, m_it, m_stage, commits, accfull,
                             accempty, acc, drained, epi))
        # ---- a TMA transfer completes (complete_tx on the stage's full barrier)
        for it in tma:
            sl = it % STAGES
            succ.append((p_it, tma - {it}, rep(stage, sl, ("landed", it)), consumed, rep(full, sl, full[sl] + 1), empty, m_it, m_stage, commits,
                         accfull, accempty, acc, drained, epi))
        # ---- MMA thread
        if m_it < total:
            t, kb = divmod(m_it, NKB)
            a, aph = t & 1, (t >> 1) & 1
            if kb == 0 and m_stage == 0:
                if not mma_waits_acc or test(accempty[a][0], aph ^ 1):
                    if t >= 2 and any(d != t - 2 for d in drained[a]):
                        return "violation: MMAs overwrite an accumulator that is still being read", (t, drained[a])
                    succ.append((p_it, tma, stage, consumed, full, empty, m_it, 1, commits, accfull, accempty, acc, drained, epi))
            else:
                sl, ph = m_it % STAGES, (m_it // STAGES) & 1
                if test(full[sl], ph):
                    if stage[sl] != ("landed", m_it):
                        return "violation: MMA reads a stage that does not hold its k-block", (m_it, stage[sl])
                    nc = commits + (("empty", sl, m_it),)
                    if kb == NKB - 1:
                        nc = nc + (("accfull", a, t),)
                    succ.append((p_it, tma, stage, consumed, full, empty, m_it + 1, 0 if kb == NKB - 1 else 1, nc, accfull, accempty, acc, drained, epi))
        # ---- the oldest pending tcgen05.commit arrives
        if commits:
            kind, i, v = commits[0]
            if kind == "empty":
                succ.append((p_it, tma, stage, rep(consumed, i, v), full, rep(empty, i, empty[i] + 1), m_it, m_stage, commits[1:], accfull, accempty,
                             acc, drained, epi))
            else:
                succ.append((p_it, tma, stage, consumed, full, empty, m_it, m_stage, commits[1:], rep(accfull, i, accfull[i] + 1), accempty,
                             rep(acc, i, v), drained, epi))
        # ---- epilogue warps: mbar_wait(tmem_full_bar(acc), aph); tcgen05.ld ...; arrive(tmem_empty_bar(acc))
        for w in range(EW):
            t = epi[w]
            if t < NT:
                a, aph = t & 1, (t >> 1) & 1
                if test(accfull[a], aph):
                    if acc[a] != t:
                        return "violation: epilogue reads an accumulator that is not its finished tile", (w, t, acc[a])
                    c, p = accempty[a]
                    p -= 1
                    if p == 0:
                        c, p = c + 1, EW
                    succ.append((p_it, tma, stage, consumed, full, empty, m_it, m_stage, commits, accfull, rep(accempty, a, (c, p)), acc,
                                 rep(drained, a, rep(drained[a], w, t)), rep(epi, w, t + 1)))
        if not succ:
            if not (p_it == total and m_it == total and not commits and not tma and all(e == NT for e in epi)):
                return "deadlock", s
        stack.extend(succ)
    return "safe", n


def test_gemm_pipeline_is_safe_under_every_interleaving():
    # (STAGES, k-blocks per tile, tiles, epilogue warps); the kernel runs 3-6 stages, 1-120 k-blocks, 4 or 8 epilogue warps.
    # (6, 4, 6, 8) -- the kernel's own geometry -- is 3.76 M states (safe; ~90 s, not run here)
    for cfg in ((2, 2, 4, 2), (3, 2, 3, 2), (2, 3, 5, 1), (2, 1, 6, 3), (3, 3, 6, 4), (4, 4, 8, 2), (6, 30, 3, 4), (3, 2, 8, 8)):
        kind, states = explore(*cfg)
        assert kind == "safe", (cfg, kind, states)


def test_the_explorer_rejects_broken_protocols():
    assert explore(2, 2, 4, 2, producer_waits=False)[0].startswith("violation")
    assert explore(2, 2, 5, 2, mma_waits_acc=False)[0].startswith("violation")


def test_model_matches_the_kernel_source():
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pcrcg_b200", "csrc", "gemm_tc.cu")).read()
    assert "mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1);" in src
    assert "mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), Cfg::EPI_WARPS);" in src
    assert re.search(r"mbar_wait\(empty_bar\(s\), ph \^ 1u\);", src) and re.search(r"mbar_wait\(full_bar\(s\), ph\);", src)
    assert re.search(r"mbar_wait\(tmem_empty_bar\(acc\), aph \^ 1u\);", src) and re.search(r"mbar_wait\(tmem_full_bar\(acc\), aph\);", src)
    assert src.index("umma_commit(empty_bar(s));") < src.index("umma_commit(tmem_full_bar(acc));")
