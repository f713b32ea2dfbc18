"""Pure-Python .skf decoder/encoder used ONLY by the tests (oracle side).

TEST INFRASTRUCTURE -- independent of the product's C++ codec (ska.rust_b200/host/skf_codec.cpp) so
that each can check the other.  Format (reference writer
src/merge_ska_array.rs:191-204, struct field order :108-126): a Snappy *framing* stream whose
concatenated payload is one CBOR map {k, rc, names, split_kmers, variants{v,dim,data},
variant_count, ska_version, k_bits}.
"""
import struct


# ---------------------------------------------------------------- CRC-32C (Castagnoli), masked
def _make_crc_table():
    tbl = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tbl.append(c)
    return tbl


_CRC_TBL = _make_crc_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TBL[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------- Snappy raw block + framing
def snappy_raw_decompress(buf: bytes) -> bytes:
    pos = 0
    ulen = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        ulen |= (b & 0x7F) << shift
        if not b & 0x80:
            break
        shift += 7
    out = bytearray()
    n = len(buf)
    while pos < n:
        tag = buf[pos]
        pos += 1
        t = tag & 3
        if t == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
        else:
            if t == 1:
                ln = 4 + ((tag >> 2) & 7)
                off = ((tag >> 5) << 8) | buf[pos]
                pos += 1
            elif t == 2:
                ln = (tag >> 2) + 1
                off = buf[pos] | (buf[pos + 1] << 8)
                pos += 2
            else:
                ln = (tag >> 2) + 1
                off = int.from_bytes(buf[pos:pos + 4], "little")
                pos += 4
            if off == 0 or off > len(out):
                raise ValueError("bad snappy copy offset")
            for _ in range(ln):  # byte-wise: copies may overlap their own output
                out.append(out[-off])
    if len(out) != ulen:
        raise ValueError("snappy length mismatch")
    return bytes(out)


def snappy_frame_decode(data: bytes) -> bytes:
    pos = 0
    out = bytearray()
    seen_id = False
    while pos < len(data):
        ctype = data[pos]
        ln = data[pos + 1] | (data[pos + 2] << 8) | (data[pos + 3] << 16)
        body = data[pos + 4:pos + 4 + ln]
        pos += 4 + ln
        if ctype == 0xFF:
            if body != b"sNaPpY":
                raise ValueError("bad stream identifier")
            seen_id = True
        elif ctype in (0x00, 0x01):
            if not seen_id:
                raise ValueError("missing stream identifier")
            crc = struct.unpack("<I", body[:4])[0]
            payload = snappy_raw_decompress(body[4:]) if ctype == 0x00 else bytes(body[4:])
            if masked_crc(payload) != crc:
                raise ValueError("crc mismatch")
            out += payload
        elif ctype == 0xFE or 0x80 <= ctype <= 0xFD:
            continue
        else:
            raise ValueError(f"unskippable chunk type {ctype:#x}")
    return bytes(out)


def snappy_frame_encode_raw(payload: bytes) -> bytes:
    """Uncompressed (type 0x01) chunks only -- valid framing, reference-loadable."""
    out = bytearray(b"\xff\x06\x00\x00sNaPpY")
    for i in range(0, len(payload), 65536):
        chunk = payload[i:i + 65536]
        ln = len(chunk) + 4
        out += bytes([0x01, ln & 0xFF, (ln >> 8) & 0xFF, (ln >> 16) & 0xFF])
        out += struct.pack("<I", masked_crc(chunk)) + chunk
    return bytes(out)


# ---------------------------------------------------------------- CBOR subset
class _Cbor:
    def __init__(self, b):
        self.b = b
        self.p = 0

    def head(self):
        ib = self.b[self.p]
        self.p += 1
        major, info = ib >> 5, ib & 31
        if info < 24:
            return major, info
        nb = {24: 1, 25: 2, 26: 4, 27: 8}.get(info)
        if nb is None:
            raise ValueError("indefinite/reserved CBOR head not supported")
        v = int.from_bytes(self.b[self.p:self.p + nb], "big")
        self.p += nb
        return major, v

    def item(self):
        major, v = self.head()
        if major == 0:
            return v
        if major == 2:
            s = self.b[self.p:self.p + v]
            self.p += v
            return bytes(s)
        if major == 3:
            s = self.b[self.p:self.p + v]
            self.p += v
            return s.decode("utf-8")
        if major == 4:
            return [self.item() for _ in range(v)]
        if major == 5:
            d = {}
            for _ in range(v):
                k = self.item()
                d[k] = self.item()
            return d
        if major == 6:
            inner = self.item()
            if v == 2:
                return int.from_bytes(inner, "big")
            return inner
        if major == 7:
            if v == 20:
                return False
            if v == 21:
                return True
            if v == 22:
                return None
        raise ValueError(f"unsupported CBOR major {major} value {v}")


def _enc_head(major, v):
    if v < 24:
        return bytes([(major << 5) | v])
    for info, nb in ((24, 1), (25, 2), (26, 4), (27, 8)):
        if v < (1 << (8 * nb)):
            return bytes([(major << 5) | info]) + v.to_bytes(nb, "big")
    raise ValueError("too large")


def _enc_uint(v):
    if v < (1 << 64):
        return _enc_head(0, v)
    raw = v.to_bytes((v.bit_length() + 7) // 8, "big")
    return b"\xc2" + _enc_head(2, len(raw)) + raw


def _enc_text(s):
    b = s.encode("utf-8")
    return _enc_head(3, len(b)) + b


def load_skf_bytes(data: bytes) -> dict:
    """-> dict(k, rc, names, split_kmers[list[int]], variants[bytes row-major], n_rows, n_samples,
              variant_count[list[int]], ska_version, k_bits)"""
    d = _Cbor(snappy_frame_decode(data)).item()
    r, n = d["variants"]["dim"]
    if d["variants"]["v"] != 1:
        raise ValueError("unexpected ndarray serde version")
    return {
        "k": d["k"], "rc": d["rc"], "names": d["names"], "split_kmers": d["split_kmers"],
        "variants": bytes(d["variants"]["data"]), "n_rows": r, "n_samples": n,
        "variant_count": d["variant_count"], "ska_version": d["ska_version"], "k_bits": d["k_bits"],
    }


def load_skf(path: str) -> dict:
    with open(path, "rb") as fh:
        return load_skf_bytes(fh.read())


def dump_skf_bytes(k, rc, names, split_kmers, variants: bytes, n_rows, n_samples, variant_count,
                   ska_version="0.5.2", k_bits=64) -> bytes:
    out = bytearray(b"\xa8")
    out += _enc_text("k") + _enc_head(0, k)
    out += _enc_text("rc") + (b"\xf5" if rc else b"\xf4")
    out += _enc_text("names") + _enc_head(4, len(names)) + b"".join(_enc_text(s) for s in names)
    out += _enc_text("split_kmers") + _enc_head(4, len(split_kmers)) + b"".join(_enc_uint(v) for v in split_kmers)
    out += _enc_text("variants") + b"\xa3" + _enc_text("v") + b"\x01"
    out += _enc_text("dim") + b"\x82" + _enc_head(0, n_rows) + _enc_head(0, n_samples)
    out += _enc_text("data") + _enc_head(4, len(variants)) + b"".join(_enc_head(0, v) for v in variants)
    out += _enc_text("variant_count") + _enc_head(4, len(variant_count)) + b"".join(_enc_head(0, v) for v in variant_count)
    out += _enc_text("ska_version") + _enc_text(ska_version)
    out += _enc_text("k_bits") + _enc_head(0, k_bits)
    return snappy_frame_encode_raw(bytes(out))
