"""The CUDA selection kernel does not walk tiles sequentially like PixelSelector::select
(PixelSelector2.cpp:290-433); it applies three order-free rules derived from that walk
(DESIGN.md §4).  This CPU test checks the derivation itself: a numpy implementation of the rules
must reproduce the oracle's literal sequential restatement, including exact ties (quantised
gradients make level-1/2 values tie across the 2x2 / 4x4 pixels that share a coarse pixel)."""
import numpy as np
import pytest


def rule_based_select(g0, g1, g2, thsS, w, h, pot):
    """-> status map (0/1/2/4) and (n2, n3, n4) from the three rules."""
    out = np.zeros(w * h, np.uint8)
    w1, w2, w32 = w // 2, w // 4, w // 32
    n = [0, 0, 0]
    f32 = np.float32
    for y4 in range(0, h, 4 * pot):
        for x4 in range(0, w, 4 * pot):
            tile_any01 = False
            best4, idx4 = f32(0), -1
            for q in range(4):                              # 2pot tiles, row-major
                x34, y34 = x4 + (q & 1) * 2 * pot, y4 + (q >> 1) * 2 * pot
                if x34 >= w or y34 >= h:
                    continue
                any0 = False
                best3, idx3 = f32(0), -1
                for b in range(4):                          # pot blocks, row-major
                    x0, y0 = x34 + (b & 1) * pot, y34 + (b >> 1) * pot
                    if x0 >= w or y0 >= h:
                        continue
                    best2, idx2 = f32(0), -1
                    for yf in range(y0, min(y0 + pot, h)):
                        for xf in range(x0, min(x0 + pot, w)):
                            if xf < 4 or xf >= w - 5 or yf < 4 or yf > h - 4:
                                continue
                            idx = xf + w * yf
                            th0 = thsS[(xf >> 5) + (yf >> 5) * w32]
                            th1 = f32(th0 * f32(0.75))
                            th2 = f32(th1 * f32(0.5625))
                            a0, a1, a2 = g0[idx], g1[(xf >> 1) + (yf >> 1) * w1], g2[(xf >> 2) + (yf >> 2) * w2]
                            if a0 > th0:
                                any0 = True
                                if a0 > best2:
                                    best2, idx2 = a0, idx
                            if a1 > th1:
                                tile_any01 = True
                                if a1 > best3:
                                    best3, idx3 = a1, idx
                            if a2 > th2 and a2 > best4:
                                best4, idx4 = a2, idx
                    if idx2 > 0:
                        out[idx2] = 1
                        n[0] += 1
                if any0:
                    tile_any01 = True
                elif idx3 > 0:
                    out[idx3] = 2
                    n[1] += 1
            if not tile_any01 and idx4 > 0:
                out[idx4] = 4
                n[2] += 1
    return out.reshape(h, w), n


@pytest.mark.parametrize("seed,quant", [(0, 1), (1, 16), (2, 64)])
def test_rules_reproduce_sequential_select(oracle_plain, tum_calib, seed, quant):
    w, h = 160, 128
    rng = np.random.default_rng(seed)
    # coarse quantisation of a smooth random field -> plateaus, exact ties and many sub-threshold tiles
    base = rng.normal(size=(h // 8 + 2, w // 8 + 2))
    img = np.kron(base, np.ones((8, 8)))[:h, :w] * 40 + 128 + rng.normal(0, 6, (h, w))
    img = (np.clip(img, 0, 255) // quant * quant).astype(np.uint8)
    bgr = np.repeat(img[:, :, None], 3, axis=2)
    st = oracle_plain.stages(bgr)
    p = oracle_plain.default_params()
    p.num_want = 10 ** 9 if False else 3000
    hdl = oracle_plain.create(tum_calib, p)
    oracle_plain.set_frame(hdl, 0, bgr, np.ones((h, w), np.uint16))
    m_ref, info = oracle_plain.get_selection_debug(hdl, 0, w, h)
    oracle_plain.destroy(hdl)
    g0, g1, g2 = (x.reshape(-1) for x in st["g2"])
    m_rule, n = rule_based_select(g0, g1, g2, st["ths_smoothed"].reshape(-1), w, h, info["pot"])
    assert [info["n2"], info["n3"], info["n4"]] == n
    # the oracle's map is after sub-sampling: every kept pixel must be a rule-selected pixel of the same
    # level, and without sub-sampling (quotia >= 0.95) the maps are identical
    kept = m_ref != 0
    assert np.array_equal(m_ref[kept], m_rule[kept])
    if sum(n) <= 3000 / 0.95:
        assert np.array_equal(m_ref, m_rule)
    assert sum(n) > 50
