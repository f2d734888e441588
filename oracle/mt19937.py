"""TEST INFRASTRUCTURE ONLY -- CPU restatement of numpy's legacy global generator (np.random.rand), which is what
GriffinLim.forward draws its initial phase from (fairseq/models/text_to_speech/vocoder.py:103:
``np.random.rand(*specgram.shape)``).

Third-party algorithm: numpy's legacy ``RandomState`` = MT19937 (Matsumoto & Nishimura 1998; numpy/random/_mt19937.pyx,
randomkit ``rk_double``).  Pinned against numpy itself (tests/test_oracle_golden.py): the state tuple is
``('MT19937', key[624] uint32, pos, has_gauss, cached_gaussian)``; output k is ``temper(key[pos])`` with the whole block
regenerated when ``pos == 624``; a double is ``((a >> 5) * 2**26 + (b >> 6)) / 2**53`` from two consecutive outputs.

Written as ONE infinite untempered sequence X with X[0:624] = key and
    X[n] = X[n - 227] ^ twist(X[n - 624], X[n - 623])        (mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1])
of which the outputs are temper(X[pos]), temper(X[pos + 1]), ...  The closest dependency is 227 back (``advance`` walks
the sequence 227 elements at a time, the device kernel 454 with a two-step chain).  ``next_block`` is the same recurrence
with the dependencies inside a block substituted away (all 624 words from the previous block in one parallel step): a
cross-check here, and a kernel variant that was measured and lost (three times the twists: 0.42 vs 0.22 ms).
"""
import numpy as np


def twist(u, v):
    y = (u & np.uint32(0x80000000)) | (v & np.uint32(0x7FFFFFFF))
    return (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908B0DF), np.uint32(0))


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


def next_block(old):
    """X[624 (b + 1) : 624 (b + 2)] from X[624 b : 624 (b + 1)], every word from the OLD block alone."""
    old = np.asarray(old, np.uint32)
    new = np.empty(624, np.uint32)
    i = np.arange(0, 227)
    new[i] = old[i + 397] ^ twist(old[i], old[i + 1])
    i = np.arange(227, 454)
    new[i] = old[i + 170] ^ twist(old[i - 227], old[i - 226]) ^ twist(old[i], old[i + 1])
    i = np.arange(454, 624)
    nxt = np.append(old[455:624], new[0])  # old[624] := new[0]
    new[i] = old[i - 57] ^ twist(old[i - 454], old[i - 453]) ^ twist(old[i - 227], old[i - 226]) ^ twist(old[i], nxt)
    return new


def advance(key, pos, n_words):
    """(raw words X[pos : pos + n_words], new key, new pos) -- what consuming n_words 32-bit outputs does to the state."""
    key = np.asarray(key, np.uint32)
    end = pos + n_words
    q = max((end - 1) // 624, 0)
    n_end = 624 * (q + 1)
    x = np.empty(max(n_end, 624), np.uint32)
    x[:624] = key
    for base in range(624, n_end, 227):
        n = np.arange(base, min(base + 227, n_end))
        x[n] = x[n - 227] ^ twist(x[n - 624], x[n - 623])
    return x[pos:end].copy(), x[624 * q: 624 * q + 624].copy(), end - 624 * q


def rand(state, shape):
    """np.random.rand(*shape) from the legacy state tuple ``state``; returns (uniforms float64, new state tuple)."""
    name, key, pos, has_gauss, cached = state
    assert name == "MT19937"
    n = int(np.prod(shape))
    words, key2, pos2 = advance(key, pos, 2 * n)
    w = temper(words)
    a, b = w[0::2] >> np.uint32(5), w[1::2] >> np.uint32(6)
    u = (a.astype(np.float64) * 67108864.0 + b.astype(np.float64)) / 9007199254740992.0
    return u.reshape(shape), (name, key2, pos2, has_gauss, cached)
