"""TEST INFRASTRUCTURE — CPU restatement of the reference's heuristic baselines (heuristic.py:11-577) as placement-
selection rules over an environment view, used to check the batched CUDA heuristics kernel.  Not part of the product.

Every baseline is "enumerate placements in a fixed order, ask Space.drop_box_virtual for feasibility (+ rest height or the
updated height map), score, keep the best under a strict comparison" — so each is restated as (candidate order, score,
comparison).  The view needs: ems() (list order), drop_box_virtual(dims, lx, ly) -> (feasible, rest_height), plain(),
next_box, container (W, L, H), setting.  `OracleDiscrete` (pct_oracle.py) provides all of these.

    LSAH       heuristic.py:138-226   EMS x rot at the EMS origin, least surface area of the packed hull, tie -> tighter EMS
    OnlineBPH  heuristic.py:364-424   EMS sorted deep-bottom-left x rot, first feasible (no fit test)
    BR         heuristic.py:500-577   EMS x rot, score of the EMS itself (volume + #item types that fit, +10 if all)
    MACS       heuristic.py:11-131    EMS x rot x 4 corners, maximal usable space left below the rest height
    DBL        heuristic.py:431-493   grid x rot, min lx + ly + 100*rest_height
    HM         heuristic.py:232-293   grid x rot, min lx + ly + 100*sum(height map after placement)
    RANDOM     heuristic.py:300-357   grid x rot, uniform over the feasible ones (counter-based generator here; the
                                       reference draws from the global numpy RNG, so only the candidate set is comparable)
"""
import numpy as np

from pct_oracle import rnd_u64

NAMES = ("LSAH", "OnlineBPH", "BR", "MACS", "DBL", "HM", "RANDOM")
# `x, y, z = next_box` / `y, x, z = ...` / `z, x, y = ...` / `z, y, x = ...` / `x, z, y = ...` / `y, z, x = ...`
# (heuristic.py:176-187): index of next_box that becomes (x, y, z).  NOT the EMSPoint rotation table.
ROT = ((0, 1, 2), (1, 0, 2), (1, 2, 0), (2, 1, 0), (0, 2, 1), (2, 0, 1))
BAD_ROW = np.array([1, 0, 0, 1, 0, 0, 0, 0, 1.0])  # "no placement": a non-zero leaf row of extent 0 cannot match any item -> the env ends the episode


def bad_row(container, continuous):
    """The continuous LeafNode2Action (C:bin3D.py:151-167) never raises: an unmatched row just picks next_box[0] as z, so the
    "no placement" row there has to fail Space.drop_box's bounds test instead (lx = W + 1 > W, C:space.py:336)."""
    if not continuous:
        return BAD_ROW.copy()
    return np.array([container[0] + 1.0, 0, 0, container[0] + 1.0, 0, 0, 0, 0, 1.0])


def rot_dims(nb, rot):
    return [nb[ROT[rot][0]], nb[ROT[rot][1]], nb[ROT[rot][2]]]


def fresh_state(container):
    """LSAH's running footprint of the packed items (heuristic.py:146-147)"""
    return {"max": [0, 0], "min": [container[0], container[1]]}


def level_max_empty(free):
    """heuristic.py:17-43 for one height level: `free` is the (W, L) boolean map of empty cells"""
    W, L = free.shape
    hist = np.zeros((W, L), dtype=np.int64)
    for i in range(W - 1, -1, -1):
        for j in range(L):
            if i == W - 1:
                hist[i, j] = int(free[i, j])
            elif free[i, j]:
                hist[i, j] = hist[i + 1, j] + 1
    best = 0
    for i in range(W):
        for j in range(L):
            v = hist[i, j]
            if v == 0 or (j > 0 and v == hist[i, j - 1]):
                continue
            j2 = j
            while j2 != L - 1 and not hist[i, j2 + 1] < v:
                j2 += 1
            j1 = j
            while j1 != 0 and not hist[i, j1 - 1] < v:
                j1 -= 1
            best = max(best, int(v) * (j2 - j1 + 1))
    return best


def macs_score(plain, dims, lx, ly, h):
    """calc_maximal_usable_spaces(updated_container, h): a voxel (i, j, k) of the tracked container is non-zero exactly when
    k < height map after the placement (boxes + the space under them, heuristic.py:47-52)"""
    hm = plain.copy()
    hm[lx:lx + dims[0], ly:ly + dims[1]] = h + dims[2]
    return sum(level_max_empty(hm <= k) for k in range(h))


def choose(name, view, state, item_set=None, seed=0, gid=0, t=0):
    """-> (dims, lx, ly) of the selected placement, or None when the baseline finds no feasible one (episode over)"""
    nb = view.next_box
    W, L, H = view.container
    R = 6 if view.setting == 2 else 2
    best, best_score = None, None
    if name in ("LSAH", "OnlineBPH", "BR", "MACS"):
        ems = [list(e) for e in view.ems()]
        if name == "OnlineBPH":
            ems = sorted(ems, key=lambda e: (e[2], e[1], e[0]))
        best_ext = None
        best_score = {"LSAH": W * L + L * H + H * W, "BR": -1e10, "MACS": -1e10}.get(name)
        for e in ems:
            ext = (e[3] - e[0], e[4] - e[1], e[5] - e[2])
            for rot in range(R):
                d = rot_dims(nb, rot)
                if name == "OnlineBPH":
                    if view.drop_box_virtual(d, e[0], e[1])[0]:
                        return d, e[0], e[1]
                    continue
                if not (ext[0] >= d[0] and ext[1] >= d[1] and ext[2] >= d[2]):
                    continue
                corners = ((e[0], e[1]),) if name != "MACS" else ((e[0], e[1]), (e[3] - d[0], e[1]), (e[0], e[4] - d[1]), (e[3] - d[0], e[4] - d[1]))
                for lx, ly in corners:
                    ok, h = view.drop_box_virtual(d, lx, ly)
                    if not ok:
                        continue
                    if name == "LSAH":
                        ex = max(lx + d[0], state["max"][0]) - min(lx, state["min"][0])
                        ey = max(ly + d[1], state["max"][1]) - min(ly, state["min"][1])
                        score = ex * ey + (h + d[2]) * ey + (h + d[2]) * ex
                        if score < best_score:
                            best_score, best, best_ext = score, (d, lx, ly), ext
                        elif score == best_score and best is not None:
                            # the slack of the incumbent's EMS is measured with the CURRENT orientation's dims (:211-212)
                            if min(ext[0] - d[0], ext[1] - d[1], ext[2] - d[2]) < min(best_ext[0] - d[0], best_ext[1] - d[1], best_ext[2] - d[2]):
                                best, best_ext = (d, lx, ly), ext
                    elif name == "BR":
                        fits = sum(1 for b in item_set if ext[0] >= b[0] and ext[1] >= b[1] and ext[2] >= b[2])
                        score = ext[0] * ext[1] * ext[2] + fits + (10 if fits == len(item_set) else 0)
                        if score > best_score:
                            best_score, best = score, (d, lx, ly)
                    else:
                        score = macs_score(view.plain(), d, lx, ly, h)
                        if score > best_score:
                            best_score, best = score, (d, lx, ly)
        return best
    feasible = []
    best_score = 1e10
    plain = view.plain() if name == "HM" else None
    for lx in range(W - nb[0] + 1):  # the loop bounds use the UNROTATED item (heuristic.py:253-254)
        for ly in range(L - nb[1] + 1):
            for rot in range(R):
                d = rot_dims(nb, rot)
                ok, h = view.drop_box_virtual(d, lx, ly)
                if not ok:
                    continue
                if name == "RANDOM":
                    feasible.append((d, lx, ly))
                    continue
                if name == "DBL":
                    score = lx + ly + 100 * h
                else:  # update_height_graph (D:space.py:316-326): the footprint becomes rest height + z
                    foot = plain[lx:lx + d[0], ly:ly + d[1]]
                    score = lx + ly + 100 * (int(plain.sum()) - int(foot.sum()) + foot.size * (h + d[2]))
                if score < best_score:
                    best_score, best = score, (d, lx, ly)
    if name == "RANDOM":
        return feasible[rnd_u64(seed, gid, t) % len(feasible)] if feasible else None
    return best


def action_row(choice, container=None, continuous=False):
    if choice is None:
        return bad_row(container, continuous)
    d, lx, ly = choice
    return np.array([lx, ly, 0, lx + d[0], ly + d[1], 0, 0, 0, 1.0])


def note_placement(state, choice):
    """LSAH footprint update after the chosen placement (heuristic.py:217-220)"""
    d, lx, ly = choice
    state["max"][0] = max(state["max"][0], lx + d[0])
    state["max"][1] = max(state["max"][1], ly + d[1])
    state["min"][0] = min(state["min"][0], lx)
    state["min"][1] = min(state["min"][1], ly)


def run_episodes(name, env, episodes, item_set=None, seed=0, gid=0):
    """sequential baseline loop on one env (the shape of every function in heuristic.py): -> [(ratio, length, packed), ...]
    Works on OracleDiscrete and — LSAH / OnlineBPH / BR only, like tools.py:217-218 — on OracleContinuous (float64 geometry:
    every score below is written with the reference's operand order, so the float results are the reference's)."""
    continuous = not isinstance(env.container[0], int)
    if continuous and name not in ("LSAH", "OnlineBPH", "BR"):
        raise ValueError("only LSAH, OnlineBPH, and BR allowed for continuous environment")  # tools.py:218
    out, t = [], 0
    env.reset()
    state = fresh_state(env.container)
    while len(out) < episodes:
        c = choose(name, env, state, item_set, seed, gid, t)
        t += 1
        items = env.packed
        _, _, done, info = env.step(action_row(c, env.container, continuous))
        if c is not None:
            note_placement(state, c)
        if done:
            out.append((float(info["ratio"]), int(info["counter"]), items))
            env.reset()
            state = fresh_state(env.container)
    return out
