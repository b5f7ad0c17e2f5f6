"""Minimal reader for R's RDX2 (.rda / save()) serialisation, no R needed.

Used ONLY to turn the reference's packaged fixtures
(/root/reference/data/covidcurve_modfit.rda, covidcurve_simdata.rda) into the
small committed golden files under tests/golden/ (see tests/golden/make_golden.py).
The format walk follows SURVEY.md section 8c / Appendix A.
"""
import bz2
import gzip
import struct


class RDX:
    def __init__(self, path):
        raw = open(path, "rb").read()
        if raw[:3] == b"BZh":
            raw = bz2.decompress(raw)
        elif raw[:2] == b"\x1f\x8b":
            raw = gzip.decompress(raw)
        self.b = raw
        assert self.b[:7] == b"RDX2\nX\n", "not an RDX2 xdr stream"
        self.p = 7
        self.refs = []
        self.hdr = (self.i(), self.i(), self.i())
        self.top = self.item()

    def i(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def vec(self, n, c):
        v = list(struct.unpack_from(">%d%s" % (n, c), self.b, self.p))
        self.p += n * (8 if c == "d" else 4)
        return v

    def ln(self):
        n = self.i()
        return ((self.i() << 32) + self.i()) if n == -1 else n

    def item(self):
        return self.item_fl(self.i())

    def item_fl(self, fl):
        t, has_attr, has_tag = fl & 0xFF, bool(fl & 512), bool(fl & 1024)
        if t == 254:
            return None
        if t in (253, 242, 241, 250, 252, 251):
            return {"special": t}
        if t == 255:
            return self.refs[((fl >> 8) or self.i()) - 1]
        if t == 1:
            o = {"sym": self.item()}
            self.refs.append(o)
            return o
        if t in (249, 248, 247):
            self.i()
            o = {"ns": [self.item() for _ in range(self.i())]}
            self.refs.append(o)
            return o
        if t == 4:
            self.i()
            o = {"env": None}
            self.refs.append(o)
            _encl, frame, hashtab, _attr = self.item(), self.item(), self.item(), self.item()
            o["env"] = {"frame": frame, "hash": hashtab}
            return o
        if t in (2, 6, 3, 5):
            cells = []
            while True:
                _attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                if t != 2:
                    return {"type": t, "tag": tag, "car": car, "cdr": self.item()}
                cells.append((tag["sym"] if tag else None, car))
                fl = self.i()
                if fl & 0xFF != 2:
                    self.item_fl(fl)
                    return {"pairlist": cells}
                has_attr, has_tag = bool(fl & 512), bool(fl & 1024)
        if t == 9:
            n = self.i()
            if n == -1:
                return None
            v = self.b[self.p:self.p + n].decode("utf8", "replace")
            self.p += n
            return v
        if t in (10, 13):
            r = {"t": t, "v": self.vec(self.ln(), "i")}
        elif t == 14:
            r = {"t": t, "v": self.vec(self.ln(), "d")}
        elif t in (16, 19, 20):
            r = {"t": t, "v": [self.item() for _ in range(self.ln())]}
        elif t == 24:
            n = self.ln()
            r = {"t": t, "v": self.b[self.p:self.p + n]}
            self.p += n
        elif t == 21:
            self.i()
            r = {"t": t, "bc": self.bc1()}
        elif t in (7, 8):
            n = self.i()
            r = {"builtin": self.b[self.p:self.p + n].decode()}
            self.p += n
        elif t == 22:
            r = {"extptr": 1}
            self.refs.append(r)
            self.item()
            self.item()
        elif t == 25:
            r = {"t": t}
        elif t == 238:
            return {"altrep": (self.item(), self.item(), self.item())}
        else:
            raise ValueError("unhandled SEXP type %d at byte %d" % (t, self.p))
        if has_attr:
            r["attr"] = dict(self.item()["pairlist"])
        return r

    def bc1(self):
        self.item()
        out = []
        for _ in range(self.i()):
            t = self.i()
            out.append(self.bc1() if t == 21 else self.bclang(t) if t in (6, 2, 244, 243, 240, 239) else self.item_fl_type(t))
        return out

    def item_fl_type(self, t):
        # constants in a byte-code pool are written as "int type" followed by a normal item
        return self.item()

    def bclang(self, t):
        if t == 243:
            return self.i()
        if t in (244, 6, 2, 240, 239):
            if t == 244:
                self.i()
                t = self.i()
            if t in (240, 239):
                self.item()
            self.item()
            self.bclang(self.i())
            self.bclang(self.i())
            return None
        return self.item()


def names(o):
    return o["attr"]["names"]["v"]


def get(o, k):
    return o["v"][names(o).index(k)]


def env_items(e):
    out = dict(e["env"]["frame"]["pairlist"]) if e["env"]["frame"] else {}
    for b in (e["env"]["hash"] or {"v": []})["v"]:
        if b:
            out.update(dict(b["pairlist"]))
    return out


def load(path):
    """Return {object name: parsed object} for an .rda file."""
    return dict(RDX(path).top["pairlist"])
