"""NumPy restatement of the counter-based reset RNG of the CUDA path.  TEST INFRASTRUCTURE ONLY.

The reference draws from the global MT19937 stream (multiagent.py:48-56); the B200 path
replaces it with Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123 v1.09 constants) keyed so that every env's stream depends only on
(seed, GLOBAL env id, episode) -- shard-invariant.  This file restates the published
Philox algorithm and the draw layout documented in include/swarm_b200.h so the tests can
check the device integers bit-for-bit (known-answer vectors below) and the derived draws.

Draw layout (must match csrc/swarm_philox.cuh):
  key     = (seed & 0xffffffff, seed >> 32)
  counter = (element index, row, global env id, (episode << 3) | stream)
  stream  0: x0 (N uniforms pairs)   1: xa0   2: burn-in actions (rows 0..9)
          3: agent noise (rows 0..10)   4: particle noise (rows 0..10)
  uniform pair : u = ((o0>>5)*2^26 + (o1>>6)) * 2^-53 ,  (o2,o3) likewise   -> (x, y) in [0,1)
  normal pair  : u1 = ((o0>>8)+1)*2^-24 in (0,1], u2 = (o1>>8)*2^-24 in [0,1)
                 rad = sqrtf(-2 logf(u1)); (n_x, n_y) = rad*(cospif(2 u2), sinpif(2 u2))   [FP32]
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_X0, STREAM_XA0, STREAM_BURN, STREAM_NOISE_A, STREAM_NOISE_X = 0, 1, 2, 3, 4

# Random123 known-answer vectors for philox4x32-10: (counter[4], key[2]) -> out[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds.  Inputs broadcastable uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = [np.asarray(a, dtype=np.uint64) & MASK
                              for a in np.broadcast_arrays(c0, c1, c2, c3, k0, k1)]
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def _words(seed, env_id, episode, stream, row, n):
    idx = np.arange(n, dtype=np.uint64)
    k0 = np.uint64(seed & 0xFFFFFFFF)
    k1 = np.uint64((seed >> 32) & 0xFFFFFFFF)
    c3 = np.uint64(((episode << 3) | stream) & 0xFFFFFFFF)
    return philox4x32_10(idx, np.uint64(row), np.uint64(env_id), c3, k0, k1)


def uniform_pairs(seed, env_id, episode, stream, n):
    o0, o1, o2, o3 = _words(seed, env_id, episode, stream, 0, n)

    def dbl(a, b):
        return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0
                + (b >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)

    return np.stack([dbl(o0, o1), dbl(o2, o3)], axis=-1)


def normal_pairs(seed, env_id, episode, stream, row, n):
    o0, o1, _, _ = _words(seed, env_id, episode, stream, row, n)
    u1 = ((o0 >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)
    u2 = (o1 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float64(2.0) * np.pi) * u2.astype(np.float64)
    nx = (rad * np.cos(ang).astype(np.float32)).astype(np.float32)
    ny = (rad * np.sin(ang).astype(np.float32)).astype(np.float32)
    return np.stack([nx, ny], axis=-1).astype(np.float64)


def reset_draws(seed, env_id, episode, n_locusts, n_agents=10, n_burn_in=10):
    """All draws of one reset, in the reference's table shapes (rows 0..10 of the noise tables)."""
    x0 = uniform_pairs(seed, env_id, episode, STREAM_X0, n_locusts)
    xa0 = uniform_pairs(seed, env_id, episode, STREAM_XA0, n_agents)
    burn = np.stack([normal_pairs(seed, env_id, episode, STREAM_BURN, r, n_agents) for r in range(n_burn_in)])
    an = np.stack([normal_pairs(seed, env_id, episode, STREAM_NOISE_A, r, n_agents) for r in range(n_burn_in + 1)])
    pn = np.stack([normal_pairs(seed, env_id, episode, STREAM_NOISE_X, r, n_locusts) for r in range(n_burn_in + 1)])
    return x0, xa0, burn, an, pn
