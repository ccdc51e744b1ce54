"""Generates tests/golden/nth_element_depth_limit_sequences.npy: sequences of 20-32 distinct values on which libstdc++'s
introselect (std::nth_element, nth = 0 and 4) runs out of its depth limit 2 * lg(n) and falls into __heap_select — the one
branch random data never reaches.  Found by hill climbing on the number of partition rounds of a Python model of
__introselect; whether a sequence really reaches the branch is checked against the restatement's counter
(oracle/lsd_oracle.c::orc_heap_select_count) by tests/test_oracle_golden.py.  Rows: [n, nth, values..., -1 padding]."""
import numpy as np


def rounds(a, nth):
    a = list(a); first, last = 0, len(a); r = 0
    depth = 2 * ((last - first).bit_length() - 1) if last - first > 0 else 0
    while last - first > 3:
        if depth == 0:
            return r, True
        depth -= 1; r += 1
        mid = first + (last - first) // 2; x, y, z = first + 1, mid, last - 1
        if a[x] < a[y]:
            m = y if a[y] < a[z] else (z if a[x] < a[z] else x)
        else:
            m = x if a[x] < a[z] else (z if a[y] < a[z] else y)
        a[first], a[m] = a[m], a[first]
        lo, hi = first + 1, last
        while True:
            while a[lo] < a[first]: lo += 1
            hi -= 1
            while a[first] < a[hi]: hi -= 1
            if not lo < hi: break
            a[lo], a[hi] = a[hi], a[lo]; lo += 1
        if lo <= nth: first = lo
        else: last = lo
    return r, False


def main():
    rng = np.random.default_rng(1)
    found = []
    for n in (32, 28, 24, 20, 31):
        for nth in (0, 4):
            for restart in range(30):
                a = rng.permutation(n).astype(float)
                sc = rounds(a, nth)
                for it in range(6000):
                    i, j = rng.integers(0, n, 2)
                    b = a.copy(); b[i], b[j] = b[j], b[i]
                    s2 = rounds(b, nth)
                    if (s2[1], s2[0]) >= (sc[1], sc[0]): a, sc = b, s2
                    if sc[1]: break
                if sc[1]:
                    found.append((n, nth, a)); break
    np.save(__file__.replace("make_nth_element_sequences.py", "nth_element_depth_limit_sequences.npy"),
            np.array([np.concatenate([[n, nth], a, np.full(32 - n, -1.0)]) for n, nth, a in found]))


if __name__ == "__main__":
    main()
