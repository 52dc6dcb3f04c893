"""Exhaustive 0-1 verification of the shared-rows median-of-25 selection used by op_median (tvl1_device.cuh,
median25_pair): two vertically adjacent 5x5 windows share 20 values; min/max pairs are discarded from the shared
working set first (forgetful selection), then each output finishes with its own 5 values.  A network of min/max
operations is correct for all inputs iff it is correct for all 0/1 inputs; 2^25 inputs are checked bit-parallel."""
import numpy as np

def cswap(w, a, b):
    lo, hi = w[a] & w[b], w[a] | w[b]
    w[a], w[b] = lo, hi

def extract_minmax(w, m):       # afterwards w[0] = min, w[1] = max of w[0..m-1]
    for i in range(0, m - 1, 2):
        cswap(w, i, i + 1)
    for i in range(2, m - 1, 2):
        cswap(w, 0, i)
        cswap(w, i + 1, 1)
    if m & 1:
        cswap(w, 0, m - 1)
        cswap(w, m - 1, 1)

def network(s, o):              # s: 20 shared values, o: 5 own values -> median of the 25
    w = list(s[:14])
    for m, nxt in zip(range(14, 8, -1), range(14, 20)):
        extract_minmax(w, m)
        w[0], w[1] = s[nxt], w[m - 1]
    extract_minmax(w, 8)
    v = w[2:8] + [o[0]]
    for m, k in zip(range(7, 4, -1), range(1, 4)):
        extract_minmax(v, m)
        v[0], v[1] = o[k], v[m - 1]
    extract_minmax(v, 4)
    a, b, c = o[4], v[2], v[3]
    return (a & b) | (a & c) | (b & c)      # median of three (0/1: majority) == max(min(a,b), min(max(a,b), c))

def main():
    n_bits = 25
    words = 1 << (n_bits - 6)
    idx = np.arange(words, dtype=np.uint64)
    lane_masks = [np.uint64(m) for m in (0xAAAAAAAAAAAAAAAA, 0xCCCCCCCCCCCCCCCC, 0xF0F0F0F0F0F0F0F0,
                                         0xFF00FF00FF00FF00, 0xFFFF0000FFFF0000, 0xFFFFFFFF00000000)]
    full = np.uint64(0xFFFFFFFFFFFFFFFF)
    x = []
    for b in range(n_bits):
        if b < 6:
            x.append(np.full(words, lane_masks[b], np.uint64))
        else:
            x.append(np.where((idx >> np.uint64(b - 6)) & np.uint64(1), full, np.uint64(0)))
    # expected: majority (>= 13 ones): popcount via per-bit counters
    cnt = np.zeros((5, words), np.uint64)   # 5-bit ripple counter per case, bit-sliced
    for v in x:
        carry = v
        for k in range(5):
            cnt[k], carry = cnt[k] ^ carry, cnt[k] & carry
    # value >= 13  <=>  bit4 | (bit3 & bit2 & (bit1 | bit0))
    expected = cnt[4] | (cnt[3] & cnt[2] & (cnt[1] | cnt[0]))
    got = network(x[:20], x[20:])
    bad = int(np.count_nonzero(got ^ expected))
    print("words with a wrong median:", bad, "of", words)
    assert bad == 0
    return 0

if __name__ == "__main__":
    raise SystemExit(main())
