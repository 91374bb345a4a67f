"""Message-contract-compatible stand-in for the reference's cognitive_engine block.

The reference block (python/cognitive_engine.py:39-149) takes `packet_info` dicts from flex_rx and emits
`configuration` dicts {modulation, inner_code, outer_code} to flex_tx, one sqlite transaction per packet.
That cannot drive thousands of channels, so this is a vectorised epsilon-greedy policy over the same 616
configurations with the same numbering: config_id = modulation*56 + inner_code*8 + outer_code + 1
(python/cognitive_engine.py:87).  It is a caller of the hot path, not part of it: numpy only."""
import numpy as np

N_MOD, N_INNER, N_OUTER = 11, 7, 8
N_CONFIGS = N_MOD * N_INNER * N_OUTER          # 616

BITS = np.array([1, 2, 3, 4, 1, 2, 3, 2, 4, 5, 6], np.float64)                  # bits/symbol by modulation index
INNER_RATE = np.array([1, 1 / 2, 2 / 3, 4 / 5, 5 / 6, 6 / 7, 7 / 8], np.float64)  # what flex_tx actually sends
OUTER_RATE = np.array([1, 1 / 2, 223 / 255, 4 / 7, 8 / 12, 16 / 22, 32 / 39, 64 / 72], np.float64)


def config_id(modulation, inner_code, outer_code):
    return modulation * 56 + inner_code * 8 + outer_code + 1


def from_config_id(cid):
    c = np.asarray(cid) - 1
    return c // 56, (c % 56) // 8, c % 8


def goodput(modulation, inner_code, outer_code, payload_valid):
    return BITS[modulation] * INNER_RATE[inner_code] * OUTER_RATE[outer_code] * payload_valid


class EpsilonGreedy(object):
    """One bandit per channel, all channels updated with array operations."""

    def __init__(self, n_channels, epsilon=0.1, seed=0):
        self.n = n_channels
        self.eps = epsilon
        self.trials = np.zeros((n_channels, N_CONFIGS), np.int64)
        self.reward = np.zeros((n_channels, N_CONFIGS), np.float64)
        self.rng = np.random.default_rng(seed)

    def update(self, channels, packet_infos):
        """packet_infos: dicts as published on flex_rx's packet_info port."""
        for ch, info in zip(channels, packet_infos):
            m, i, o = info["modulation"], info["inner_code"], info["outer_code"]
            if min(m, i, o) < 0:
                continue                                   # schemes outside the tables are not learned from
            cid = config_id(m, i, o) - 1
            self.trials[ch, cid] += 1
            self.reward[ch, cid] += goodput(m, i, o, info["payload_valid"])

    def choose(self):
        """Returns per-channel `configuration` dicts for flex_tx's configuration port."""
        mean = np.where(self.trials > 0, self.reward / np.maximum(self.trials, 1), np.inf)   # untried first
        best = mean.argmax(axis=1)
        explore = self.rng.random(self.n) < self.eps
        pick = np.where(explore, self.rng.integers(0, N_CONFIGS, self.n), best) + 1
        m, i, o = from_config_id(pick)
        return [{"modulation": int(a), "inner_code": int(b), "outer_code": int(c)} for a, b, c in zip(m, i, o)]
