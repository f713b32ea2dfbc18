"""CPU models of three bit-level tricks the CUDA kernels rely on.  They do not run device code; they pin the
*reasoning* the kernels are built on (the parity of the kernels themselves is tests/test_gpu_*.py):

  * bucket_count_kernel (csrc/merge.cu): an open-addressing table whose home slot is monotone in the key and whose
    probing never wraps yields the keys in sorted order when every key is ranked inside its run of adjacent occupied
    slots -- for any insertion order;
  * transpose_words_kernel (csrc/filter.cu): the six byte-permute selectors transpose a 4 x 4 byte block;
  * iupac_letters4 (csrc/merge.cu): two byte-permutes over a 16-letter table map four 4-bit base sets to the IUPAC
    letters of bit_encoding.rs:388-453 (set bit = 2-bit code: A=1 C=2 T=4 G=8).
"""
import numpy as np


def test_order_preserving_table_ranks_equal_sorted_order():
    rng = np.random.default_rng(1)
    cap, tail = 4096, 256
    slots = cap + tail
    for trial in range(60):
        n = int(rng.integers(1, 3000))
        width = int(rng.integers(1, 1 << 44))
        if trial % 2:
            keys = np.unique((rng.random(n) ** 4 * width).astype(np.uint64))       # crowded towards the low end
        else:
            keys = np.unique(rng.integers(0, width + 1, size=n, dtype=np.uint64))
        shift = max(0, int(width).bit_length() - 12)
        tab = [None] * slots
        overflow = False
        for i in rng.permutation(len(keys)):                                       # any insertion order
            key = int(keys[i])
            s = key >> shift
            while s < slots and tab[s] is not None:
                s += 1
            if s >= slots:
                overflow = True                                                    # the kernel then splits the item
                break
            tab[s] = key
        if overflow:
            continue
        occupied = [(key, s) for s, key in enumerate(tab) if key is not None]
        out = [None] * len(occupied)
        for i, (key, s) in enumerate(occupied):
            rank = i
            j = i
            while j > 0 and occupied[j - 1][1] + 1 == occupied[j][1]:
                rank -= occupied[j - 1][0] > key
                j -= 1
            j = i
            while j + 1 < len(occupied) and occupied[j + 1][1] == occupied[j][1] + 1:
                rank += occupied[j + 1][0] < key
                j += 1
            assert out[rank] is None
            out[rank] = key
        assert out == sorted(int(x) for x in keys)


def byte_perm(x, y, sel):
    """CUDA __byte_perm: result byte i = byte (selector nibble i) of the 8 bytes {x.0..x.3, y.0..y.3}"""
    src = [(x >> (8 * i)) & 0xFF for i in range(4)] + [(y >> (8 * i)) & 0xFF for i in range(4)]
    return sum(src[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))


def test_byte_permute_selectors_transpose_a_4x4_block():
    rng = np.random.default_rng(2)
    for _ in range(50):
        blk = rng.integers(0, 256, size=(4, 4))                                    # blk[row][sample]
        a = [sum(int(blk[r][s]) << (8 * s) for s in range(4)) for r in range(4)]   # one word per row
        t0, t1 = byte_perm(a[0], a[1], 0x5140), byte_perm(a[2], a[3], 0x5140)
        t2, t3 = byte_perm(a[0], a[1], 0x7362), byte_perm(a[2], a[3], 0x7362)
        b = [byte_perm(t0, t1, 0x5410), byte_perm(t0, t1, 0x7632), byte_perm(t2, t3, 0x5410), byte_perm(t2, t3, 0x7632)]
        for s in range(4):
            for r in range(4):
                assert (b[s] >> (8 * r)) & 0xFF == blk[r][s]                       # word per sample, byte per row


def test_iupac_letters_by_two_byte_permutes():
    letters = b"-ACMTWYHGRSVKDBN"                                                  # index = base set (A=1 C=2 T=4 G=8)
    # the same table from the definition: set of bases -> IUPAC code
    code = {frozenset("A"): "A", frozenset("C"): "C", frozenset("G"): "G", frozenset("T"): "T", frozenset("AG"): "R",
            frozenset("CT"): "Y", frozenset("CG"): "S", frozenset("AT"): "W", frozenset("GT"): "K", frozenset("AC"): "M",
            frozenset("CGT"): "B", frozenset("AGT"): "D", frozenset("ACT"): "H", frozenset("ACG"): "V", frozenset("ACGT"): "N"}
    for m in range(1, 16):
        bases = frozenset(b for b, bit in zip("ACTG", (1, 2, 4, 8)) if m & bit)
        assert chr(letters[m]) == code[bases]
    T = [int.from_bytes(letters[4 * i:4 * i + 4], "little") for i in range(4)]
    rng = np.random.default_rng(3)
    for _ in range(200):
        sets = [int(x) for x in rng.integers(0, 16, size=4)]
        v = sum(s << (8 * i) for i, s in enumerate(sets))
        sel = (v & 0x7) | ((v >> 4) & 0x70) | ((v >> 8) & 0x700) | ((v >> 12) & 0x7000)
        lo, hi = byte_perm(T[0], T[1], sel), byte_perm(T[2], T[3], sel)
        mask = ((v >> 3) & 0x01010101) * 0xFF
        word = (hi & mask) | (lo & ~mask & 0xFFFFFFFF)
        assert [(word >> (8 * i)) & 0xFF for i in range(4)] == [letters[s] for s in sets]


def test_scatter_tiles_are_claimed_exactly_once_by_the_two_launches():
    """part_scatter2_kernel<1, *, STR> (csrc/partition.cu): the common launch takes every tile whose first and last
    position lie in one sample; launch STR = 1 has one CTA per sample boundary m and takes the boundary's tile iff the
    boundary is strictly inside the tile and is the first such boundary of that tile.  Together: every tile once, and
    inside a straddling tile the per-sample rounds [a, b) tile the tile's positions."""
    rng = np.random.default_rng(7)
    T = 4096
    for trial in range(200):
        n_samples = int(rng.integers(1, 40))
        lens = rng.integers(0 if trial % 5 == 0 else 1, 3 * T, size=n_samples)
        if trial % 3 == 0:
            lens = (lens // T) * T + (rng.integers(0, 2, size=n_samples) * rng.integers(0, T, size=n_samples))   # many boundaries on tile starts
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        n_pos = int(off[-1])
        if n_pos == 0:
            continue
        n_tiles = (n_pos + T - 1) // T
        claimed = np.zeros(n_tiles, dtype=np.int64)
        sid = lambda p: int(np.searchsorted(off[:n_samples], p, side="right")) - 1      # last s with off[s] <= p
        for t in range(n_tiles):                                                          # STR = 0
            beg, end = t * T, min(n_pos, (t + 1) * T)
            hi = sid(end - 1)
            lo = hi if off[hi] <= beg else hi - 1                                         # the kernel only needs "same or not"
            assert (lo == hi) == (sid(beg) == hi)
            if lo == hi:
                claimed[t] += 1
        for m in range(1, n_samples):                                                     # STR = 1, CTA m - 1
            p = int(off[m])
            t = p // T
            if p >= n_pos or p == t * T or off[m - 1] > t * T:
                continue
            claimed[t] += 1
            beg, end = t * T, min(n_pos, (t + 1) * T)
            cnt = end - beg
            covered = np.zeros(cnt, dtype=np.int64)
            for s in range(sid(beg), sid(end - 1) + 1):
                a = min(cnt, off[s] - beg) if off[s] > beg else 0
                b = min(cnt, off[s + 1] - beg) if off[s + 1] > beg else 0
                covered[a:b] += 1
                assert all(sid(beg + i) == s for i in (a, b - 1) if a < b)
            assert (covered == 1).all()
        assert (claimed == 1).all(), (trial, off, claimed)


def test_byte_histogram_counts_exactly_the_region():
    """part_hist_bytes_kernel: a region [beg, end) of the digit-byte array is read as aligned 16-byte words; whole words
    inside the region take the unchecked path, the two ragged ends are checked byte by byte -- every byte once."""
    rng = np.random.default_rng(11)
    data = rng.integers(0, 256, size=5000, dtype=np.uint8)
    for _ in range(300):
        beg = int(rng.integers(0, 4900))
        end = int(rng.integers(beg, 5000))
        hist = np.zeros(256, dtype=np.int64)
        for w in range(beg >> 4, (end + 15) >> 4):
            p = w << 4
            word = np.zeros(16, dtype=np.uint8)
            word[:max(0, min(16, len(data) - p))] = data[p:p + 16]                          # (the array is allocated 16 bytes longer)
            if p >= beg and p + 16 <= end:
                np.add.at(hist, word, 1)
            else:
                for b in range(16):
                    if beg <= p + b < end:
                        hist[word[b]] += 1
        assert np.array_equal(hist, np.bincount(data[beg:end], minlength=256))


def test_emit_set_bit_comes_straight_from_the_record():
    """bucket_emit_smem_kernel: record = key_low << (sb + 2) | sample << 2 | middle; the set word of a row holds eight
    samples x four bits, so bit = (sample & 7) * 4 + middle = record & 31 and word = sample >> 3 = (record >> 5) & (sid_mask >> 3);
    with fewer than eight samples the key bits that follow at once are masked off."""
    rng = np.random.default_rng(13)
    for sb in range(1, 13):
        sid_mask = (1 << sb) - 1
        word_mask = sid_mask >> 3
        bit_mask = 31 if sb >= 3 else (1 << (sb + 2)) - 1
        for _ in range(2000):
            sample = int(rng.integers(0, 1 << sb)); mid = int(rng.integers(0, 4)); key_low = int(rng.integers(0, 1 << 40))
            rec = (key_low << (sb + 2)) | (sample << 2) | mid
            lo = rec & 0xFFFFFFFF
            assert (lo >> 5) & word_mask == sample >> 3
            assert lo & bit_mask == 4 * (sample & 7) + mid
