"""TEST INFRASTRUCTURE ONLY - baseline JPEG entropy decoder: file bytes -> quantised coefficient blocks per component
(zigzag order), quantisation tables and sampling factors.  Used to compare the GPU thumbnail encoder with the real
libavcodec MJPEG encoder at the coefficient level (tests/test_parity_jpeg.py)."""
import numpy as np


def parse(jpeg):
    b = jpeg
    i = 2
    qt, huff, comps, scan = {}, {}, [], None
    w = h = 0
    while i < len(b):
        assert b[i] == 0xFF, hex(b[i])
        m = b[i + 1]
        n = (b[i + 2] << 8) | b[i + 3]
        p = b[i + 4:i + 2 + n]
        if m == 0xDB:
            k = 0
            while k < len(p):
                assert p[k] >> 4 == 0
                qt[p[k] & 15] = list(p[k + 1:k + 65])
                k += 65
        elif m == 0xC0:
            h, w = (p[1] << 8) | p[2], (p[3] << 8) | p[4]
            comps = [(p[6 + 3 * c], p[7 + 3 * c] >> 4, p[7 + 3 * c] & 15, p[8 + 3 * c]) for c in range(p[5])]
        elif m == 0xC4:
            k = 0
            while k < len(p):
                tc_th, bits = p[k], list(p[k + 1:k + 17])
                vals = list(p[k + 17:k + 17 + sum(bits)])
                table, code, v = {}, 0, 0
                for ln in range(1, 17):
                    for _ in range(bits[ln - 1]):
                        table[(ln, code)] = vals[v]
                        code += 1
                        v += 1
                    code <<= 1
                huff[tc_th] = table
                k += 17 + sum(bits)
        elif m == 0xDA:
            scan = [(p[1 + 2 * c], p[2 + 2 * c] >> 4, p[2 + 2 * c] & 15) for c in range(p[0])]
            i += 2 + n
            break
        i += 2 + n
    data = bytearray()
    while i < len(b):
        if b[i] == 0xFF:
            if b[i + 1] == 0:
                data.append(0xFF)
                i += 2
                continue
            if b[i + 1] == 0xD9:
                break
            raise ValueError("unexpected marker %x" % b[i + 1])
        data.append(b[i])
        i += 1
    bits = np.unpackbits(np.frombuffer(bytes(data), np.uint8))
    pos = 0

    def getbits(n):
        nonlocal pos
        v = 0
        for _ in range(n):
            v = (v << 1) | int(bits[pos])
            pos += 1
        return v

    def decode(table):
        nonlocal pos
        code = 0
        for ln in range(1, 17):
            code = (code << 1) | int(bits[pos])
            pos += 1
            if (ln, code) in table:
                return table[(ln, code)]
        raise ValueError("bad huffman code")

    def extend(v, s):
        return v if s == 0 or v >= (1 << (s - 1)) else v - (1 << s) + 1

    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    mx, my = (w + 8 * hmax - 1) // (8 * hmax), (h + 8 * vmax - 1) // (8 * vmax)
    out = {c[0]: [] for c in comps}
    pred = {c[0]: 0 for c in comps}
    sel = {s[0]: (s[1], s[2]) for s in scan}
    for _ in range(mx * my):
        for cid, hs, vs, _tq in comps:
            td, ta = sel[cid]
            for _k in range(hs * vs):
                blk = [0] * 64
                s = decode(huff[td])
                pred[cid] += extend(getbits(s), s)
                blk[0] = pred[cid]
                k = 1
                while k < 64:
                    rs = decode(huff[0x10 | ta])
                    r, s = rs >> 4, rs & 15
                    if s == 0:
                        if r == 15:
                            k += 16
                            continue
                        break
                    k += r
                    blk[k] = extend(getbits(s), s)
                    k += 1
                out[cid].append(blk)
    return {"width": w, "height": h, "components": comps, "qt": qt, "blocks": {k: np.array(v, np.int32) for k, v in out.items()}}
