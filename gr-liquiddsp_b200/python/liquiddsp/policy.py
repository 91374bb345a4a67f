"""Message-contract-compatible stand-in for the reference's cognitive_engine block.

The reference block (python/cognitive_engine.py:39-149) takes `packet_info` dicts from flex_rx and emits
`configuration` dicts {modulation, inner_code, outer_code} to flex_tx, one sqlite transaction per packet.
That cannot drive thousands of channels, so this is a vectorised epsilon-greedy policy over the same 616
configurations with the same numbering: config_id = modulation*56 + inner_code*8 + outer_code + 1
(python/cognitive_engine.py:87).  `extended=True` adds the additive scheme indices of the blocks (QAM128/256,
v27p34, the K = 9 family: modulation 11..12, inner_code 7..14) as configurations 617.. -- flagged non-reference.
It is a caller of the hot path, not part of it: numpy only, every update and choice is an array operation."""
import numpy as np

N_MOD, N_INNER, N_OUTER = 11, 7, 8
N_CONFIGS = N_MOD * N_INNER * N_OUTER          # 616
N_MOD_EXT, N_INNER_EXT = 13, 15

BITS = np.array([1, 2, 3, 4, 1, 2, 3, 2, 4, 5, 6, 7, 8], np.float64)            # bits/symbol by modulation index
INNER_RATE = np.array([1, 1 / 2, 2 / 3, 4 / 5, 5 / 6, 6 / 7, 7 / 8,              # what flex_tx actually sends
                       3 / 4, 1 / 2, 2 / 3, 3 / 4, 4 / 5, 5 / 6, 6 / 7, 7 / 8], np.float64)
OUTER_RATE = np.array([1, 1 / 2, 223 / 255, 4 / 7, 8 / 12, 16 / 22, 32 / 39, 64 / 72], np.float64)


def config_id(modulation, inner_code, outer_code):
    return modulation * 56 + inner_code * 8 + outer_code + 1


def from_config_id(cid):
    c = np.asarray(cid) - 1
    return c // 56, (c % 56) // 8, c % 8


def _extended_tables():
    """(modulation, inner, outer) of every configuration: 1..616 as the reference numbers them, then the combinations
    that use an extension index, in (modulation, inner, outer) order."""
    m, i, o = from_config_id(np.arange(1, N_CONFIGS + 1))
    ext = [(a, b, c) for a in range(N_MOD_EXT) for b in range(N_INNER_EXT) for c in range(N_OUTER) if a >= N_MOD or b >= N_INNER]
    e = np.array(ext, np.int64)
    return np.concatenate([m, e[:, 0]]), np.concatenate([i, e[:, 1]]), np.concatenate([o, e[:, 2]])


def goodput(modulation, inner_code, outer_code, payload_valid):
    return BITS[modulation] * INNER_RATE[inner_code] * OUTER_RATE[outer_code] * payload_valid


class EpsilonGreedy(object):
    """One bandit per channel, all channels updated and queried with array operations."""

    def __init__(self, n_channels, epsilon=0.1, seed=0, extended=False):
        self.n = n_channels
        self.eps = epsilon
        if extended:
            self.cfg_m, self.cfg_i, self.cfg_o = _extended_tables()
        else:
            self.cfg_m, self.cfg_i, self.cfg_o = from_config_id(np.arange(1, N_CONFIGS + 1))
        self.n_cfg = len(self.cfg_m)
        self._lut = np.full((N_MOD_EXT, N_INNER_EXT, N_OUTER), -1, np.int64)     # (m, i, o) -> column
        self._lut[self.cfg_m, self.cfg_i, self.cfg_o] = np.arange(self.n_cfg)
        self.trials = np.zeros((n_channels, self.n_cfg), np.int64)
        self.reward = np.zeros((n_channels, self.n_cfg), np.float64)
        # mean reward per (channel, configuration), +inf while untried: kept up to date entry by entry in update_arrays, so
        # that a choice is one argmax per channel instead of a pass of divisions over the whole table
        self.mean = np.full((n_channels, self.n_cfg), np.inf, np.float64)
        self.rng = np.random.default_rng(seed)

    def recompute(self):
        """Rebuild the cached means from `trials` / `reward` (only needed after writing to those arrays directly)."""
        self.mean = np.where(self.trials > 0, self.reward / np.maximum(self.trials, 1), np.inf)

    def update_arrays(self, channels, modulation, inner_code, outer_code, payload_valid):
        """One entry per received packet (arrays of equal length); schemes outside the tables (-1) are not learned from."""
        ch = np.asarray(channels, np.int64)
        m, i, o = (np.asarray(a, np.int64) for a in (modulation, inner_code, outer_code))
        pv = np.asarray(payload_valid, np.float64)
        ok = (m >= 0) & (i >= 0) & (o >= 0) & (m < N_MOD_EXT) & (i < N_INNER_EXT) & (o < N_OUTER)
        col = np.where(ok, self._lut[np.where(ok, m, 0), np.where(ok, i, 0), np.where(ok, o, 0)], -1)
        ok &= col >= 0
        ch, col, m, i, o, pv = ch[ok], col[ok], m[ok], i[ok], o[ok], pv[ok]
        # (bincount over the flat index: the same sums as np.add.at, an order of magnitude faster)
        flat = ch * self.n_cfg + col
        uniq, inv = np.unique(flat, return_inverse=True)
        t = self.trials.reshape(-1)
        r = self.reward.reshape(-1)
        t[uniq] += np.bincount(inv, minlength=len(uniq))
        r[uniq] += np.bincount(inv, weights=goodput(m, i, o, pv), minlength=len(uniq))
        self.mean.reshape(-1)[uniq] = r[uniq] / t[uniq]

    def update(self, channels, packet_infos):
        """packet_infos: dicts as published on flex_rx's packet_info port."""
        if not len(packet_infos):
            return
        self.update_arrays(channels, [p["modulation"] for p in packet_infos], [p["inner_code"] for p in packet_infos],
                           [p["outer_code"] for p in packet_infos], [p["payload_valid"] for p in packet_infos])

    def choose_arrays(self):
        """Per-channel (modulation, inner_code, outer_code) index arrays."""
        best = self.mean.argmax(axis=1)                   # untried (+inf) first, then the best mean; first index on ties
        explore = self.rng.random(self.n) < self.eps
        pick = np.where(explore, self.rng.integers(0, self.n_cfg, self.n), best)
        return self.cfg_m[pick], self.cfg_i[pick], self.cfg_o[pick]

    def choose(self):
        """Returns per-channel `configuration` dicts for flex_tx's configuration port."""
        m, i, o = self.choose_arrays()
        return [{"modulation": int(a), "inner_code": int(b), "outer_code": int(c)} for a, b, c in zip(m, i, o)]
