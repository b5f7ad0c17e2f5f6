"""R serialisation (XDR, format version 2) without R: reader and writer, and the fit-object writer of SURVEY.md 8(f)-3.

A run on a GPU box has no R.  `write_fit_rds` turns the result of `api.run_IFRmodel_age` into an `.rds` file that
`readRDS()` opens on a machine that has COVIDCurve installed: the `mcmcout` element is a genuine `drjacoby_output`
(list(output = data.frame, diagnostics, parameters), the field list of /root/reference/R/main.R:246-270 and of the stored
fit data/covidcurve_modfit.rda), so every consumer in R/summarize_plot.R (:70,167,280,690 assert on
that class) works on it unchanged.  `inputs$IFRmodel` is an R6 object (an environment full of closures) in the reference;
closures cannot be produced without R, so it is written as a plain named list of the model's fields with class
"IFRmodel_fields", and `covidcurve_b200/r/ccb200.R: ccb200_read_fit()` rebuilds the R6 object with the public setters of
R/model.R:163-355.

The object model is the one of the reader: a dict {"t": SEXPTYPE, "v": payload, "attr": {name: object}, "obj": bool,
"lev": gp bits}; CHARSXP are Python str (None = NA_character_), NULL is None, symbols are {"sym": name}.  Pinned against
R itself: parsing the reference's data/covidcurve_simdata.rda (written by R 4.0.3) and re-serialising the parsed object
reproduces the file's byte stream exactly (tests/test_rdata.py).

Format walk: SURVEY.md Appendix A.
"""
from __future__ import annotations

import bz2
import gzip
import struct

import numpy as np

NA_INTEGER = -2147483648
NA_REAL_BITS = 0x7FF00000000007A2  # R's NA_real_: a NaN with payload 1954
_WRITER_VERSION = 0x00040003         # "written by R 4.0.3", the version that wrote the reference's fixtures
_MIN_READER = 0x00020300


# ------------------------------------------------------------------------------------------------------------ reader
class RDX:
    """Parsed stream: `.top` is the object (for an .rda: a pairlist {"pairlist": [(name, object), ...]})."""

    def __init__(self, path=None, raw=None):
        if raw is None:
            raw = open(path, "rb").read()
        if raw[:3] == b"BZh":
            raw = bz2.decompress(raw)
        elif raw[:2] == b"\x1f\x8b":
            raw = gzip.decompress(raw)
        self.b = raw
        if raw[:7] == b"RDX2\nX\n":
            self.kind, self.p = "rda", 7
        elif raw[:2] == b"X\n":
            self.kind, self.p = "rds", 2
        else:
            raise ValueError("not an XDR R serialisation stream")
        self.refs = []
        self.hdr = (self.i(), self.i(), self.i())
        if self.hdr[0] != 2:
            raise ValueError("only serialisation format version 2 is supported")
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
            n = self.ln()
            r = {"t": t, "v": np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)}
            self.p += 8 * n
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
        r["obj"], r["lev"] = bool(fl & 256), fl >> 12
        if has_attr:
            r["attr"] = dict(self.item()["pairlist"])
        return r

    def bc1(self):
        self.item()
        out = []
        for _ in range(self.i()):
            t = self.i()
            out.append(self.bc1() if t == 21 else self.bclang(t) if t in (6, 2, 244, 243, 240, 239) else self.item())
        return out

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
    """{object name: parsed object} of an .rda file."""
    return dict(RDX(path).top["pairlist"])


def read_rds(path):
    r = RDX(path)
    if r.kind != "rds":
        raise ValueError("%s is an .rda stream (use load)" % path)
    return r.top


# ------------------------------------------------------------------------------------------------------------ writer
class _Writer:
    def __init__(self):
        self.out = []
        self.syms = {}  # symbol name -> 1-based index in the reference table (R writes a REFSXP for every repeat)

    def i(self, v):
        self.out.append(struct.pack(">i", v))

    def length(self, n):
        if n > 0x7FFFFFFF:
            self.i(-1)
            self.i(n >> 32)
            self.i(n & 0xFFFFFFFF)
        else:
            self.i(n)

    def chars(self, s):
        if s is None:
            self.i(9)  # NA_STRING: no encoding bits, length -1
            self.i(-1)
            return
        b = s.encode("utf8")
        self.i(9 | ((64 if len(b) == len(s) and all(c < 128 for c in b) else 8) << 12))  # ASCII_MASK, else UTF8_MASK
        self.i(len(b))
        self.out.append(b)

    def symbol(self, name):
        if name in self.syms:
            self.i((self.syms[name] << 8) | 255)
            return
        self.i(1)
        self.chars(name)
        self.syms[name] = len(self.syms) + 1

    def pairlist(self, cells):
        for tag, car in cells:
            self.i(2 | (1024 if tag is not None else 0))
            if tag is not None:
                self.symbol(tag)
            self.item(car)
        self.i(254)

    def item(self, o):
        if o is None:
            self.i(254)
            return
        if "pairlist" in o:
            self.pairlist(o["pairlist"])
            return
        if "sym" in o:
            self.symbol(o["sym"])
            return
        t, attr = o["t"], o.get("attr")
        self.i(t | (256 if o.get("obj", bool(attr) and "class" in attr) else 0) | (512 if attr else 0) | (o.get("lev", 0) << 12))
        v = o["v"]
        self.length(len(v))
        if t in (10, 13):
            self.out.append(np.asarray(v, dtype=">i4").tobytes())
        elif t == 14:
            self.out.append(np.asarray(v, dtype=">f8").tobytes())
        elif t == 16:
            for s in v:
                self.chars(s)
        elif t == 19:
            for x in v:
                self.item(x)
        else:
            raise ValueError("cannot serialise SEXP type %d" % t)
        if attr:
            self.pairlist(list(attr.items()))


def _stream(obj, magic):
    w = _Writer()
    w.out.append(magic)
    for v in (2, _WRITER_VERSION, _MIN_READER):
        w.i(v)
    w.item(obj)
    return b"".join(w.out)


def dumps_rda(objects: dict) -> bytes:
    """Uncompressed byte stream of `save(<objects>)`."""
    return _stream({"pairlist": list(objects.items())}, b"RDX2\nX\n")


def dumps_rds(obj) -> bytes:
    """Uncompressed byte stream of `saveRDS(obj)`."""
    return _stream(obj, b"X\n")


def write_rds(path, obj, compress=True):
    raw = dumps_rds(obj)
    with open(path, "wb") as f:
        f.write(gzip.compress(raw, 6, mtime=0) if compress else raw)


# --------------------------------------------------------------------------------------------- Python -> R object model
def r_real(x):
    a = np.atleast_1d(np.asarray(x, dtype=np.float64)).reshape(-1)
    return {"t": 14, "v": a}


def r_int(x):
    return {"t": 13, "v": [int(v) for v in np.atleast_1d(np.asarray(x)).reshape(-1)]}


def r_lgl(x):
    return {"t": 10, "v": [NA_INTEGER if v is None else int(bool(v)) for v in (x if isinstance(x, (list, tuple)) else [x])]}


def r_str(x):
    return {"t": 16, "v": [None if v is None else str(v) for v in ([x] if isinstance(x, str) or x is None else list(x))]}


def r_list(items: dict, cls=None):
    o = {"t": 19, "v": [to_r(v) for v in items.values()], "attr": {"names": r_str(list(items.keys()))}}
    if cls:
        o["attr"]["class"] = r_str(cls)
    return o


def r_dataframe(df, cls=("data.frame",)):
    """data.frame with compact row names c(NA, -n); object columns become character vectors, integer columns integer,
    everything else double (the column types of the stored reference fit: chain/rung/stage chr, iteration int)."""
    cols = []
    for c in df.columns:
        s = df[c]
        if s.dtype == object or str(s.dtype).startswith(("str", "string")):
            cols.append(r_str([None if v is None or (isinstance(v, float) and np.isnan(v)) else v for v in s.tolist()]))
        elif np.issubdtype(s.dtype, np.bool_):
            cols.append({"t": 10, "v": [int(v) for v in s.tolist()]})
        elif np.issubdtype(s.dtype, np.integer):
            cols.append(r_int(s.to_numpy()))
        else:
            cols.append(r_real(s.to_numpy(dtype=np.float64)))
    return {"t": 19, "v": cols, "attr": {"names": r_str([str(c) for c in df.columns]), "class": r_str(list(cls)),
                                         "row.names": {"t": 13, "v": [NA_INTEGER, -len(df)]}}}


def to_r(x):
    """Default conversion: None -> NULL, bool -> logical, int -> integer, float / ndarray of floats -> double, str / list of
    str -> character, dict -> named list, pandas Series -> named double vector, DataFrame -> data.frame; an object that is
    already in the object model (a dict with "t") passes through."""
    import pandas as pd
    if x is None:
        return None
    if isinstance(x, dict):
        return x if "t" in x else r_list(x)
    if isinstance(x, pd.DataFrame):
        return r_dataframe(x)
    if isinstance(x, pd.Series):
        o = r_real(x.to_numpy(dtype=np.float64))
        o["attr"] = {"names": r_str([str(k) for k in x.index])}
        return o
    if isinstance(x, (bool, np.bool_)):
        return r_lgl(bool(x))
    if isinstance(x, (int, np.integer)):
        return r_int(x)
    if isinstance(x, (float, np.floating)):
        return r_real(x)
    if isinstance(x, str):
        return r_str(x)
    if isinstance(x, (list, tuple)) and all(isinstance(v, str) for v in x):
        return r_str(x)
    a = np.asarray(x)
    if a.dtype == object or a.dtype.kind in "US":
        return r_str([str(v) for v in a.reshape(-1)])
    if a.dtype.kind == "b":
        return {"t": 10, "v": [int(v) for v in a.reshape(-1)]}
    if a.dtype.kind in "iu":
        return r_int(a)
    o = r_real(a)
    if a.ndim > 1:  # R arrays are column-major
        o["v"] = np.asarray(a, dtype=np.float64).reshape(-1, order="F")
        o["attr"] = {"dim": r_int(a.shape)}
    return o


# ----------------------------------------------------------------------------------------------------- the fit object
def _model_fields(m):
    """The public fields of the IFRmodel R6 class (R/model.R:14-29) as a named list."""
    data = None
    if m.data is not None:
        data = r_list({k: r_dataframe(v, ("tbl_df", "tbl", "data.frame")) for k, v in m.data.items()})
    fields = {
        "data": data, "demog": None if m.demog is None else r_dataframe(m.demog, ("tbl_df", "tbl", "data.frame")),
        "paramdf": None if m.paramdf is None else r_dataframe(m.paramdf, ("tbl_df", "tbl", "data.frame")),
        "IFRparams": m.IFRparams, "maxMa": m.maxMa, "Knotparams": m.Knotparams, "relKnot": m.relKnot, "Infxnparams": m.Infxnparams,
        "relInfxn": m.relInfxn, "Noiseparams": m.Noiseparams, "Serotestparams": m.Serotestparams, "modparam": m.modparam,
        "sodparam": m.sodparam, "maxObsDay": m.maxObsDay, "rcensor_day": int(m.rcensor_day), "IFRdictkey": m.IFRdictkey,
    }
    o = {"t": 19, "v": [to_r(v) for v in fields.values()], "attr": {"names": r_str(list(fields.keys())), "class": r_str("IFRmodel_fields")}}
    return o


def fit_to_r(fit):
    """`IFRmodel_inf` (R/main.R:246-270) in the object model.  Numeric run settings are doubles and `rungs` an integer,
    `mc_accept` is a logical NA at rungs = 1, exactly as in the stored reference fit."""
    inp, mc = fit["inputs"], fit["mcmcout"]
    inputs = {"IFRmodel": _model_fields(inp["IFRmodel"])}
    for k in ("reparamIFR", "reparamInfxn", "reparamKnots", "account_seroreversion", "binomial_likelihood"):
        inputs[k] = r_lgl(bool(inp[k]))
    for k in ("burnin", "samples", "chains"):
        inputs[k] = r_real(float(inp[k]))
    if "rungs" in inp:
        inputs.update(rungs=r_real(float(inp["rungs"])), GTI_pow=r_real(float(inp["GTI_pow"])),
                      beta_manual=to_r(None if inp["beta_manual"] is None else np.asarray(inp["beta_manual"], dtype=np.float64)),
                      coupling_on=r_lgl(bool(inp["coupling_on"])))
    dg, pr = mc["diagnostics"], mc["parameters"]
    mc_accept = r_lgl(None) if not hasattr(dg["mc_accept"], "columns") else r_dataframe(dg["mc_accept"])
    rhat = dg["rhat"]  # unnamed double vector in df_params order, like the stored fit
    diagnostics = {"rhat": None if rhat is None else r_real(np.asarray(rhat, dtype=np.float64)), "rung_details": r_dataframe(dg["rung_details"]), "mc_accept": mc_accept}
    for k in ("move_accept", "bw"):  # extras of this implementation (not in drjacoby's list; consumers ignore them)
        if dg.get(k) is not None:
            diagnostics[k] = to_r(np.asarray(dg[k], dtype=np.float64))
    beta = pr["beta_manual"]  # drjacoby stores the ladder actually used (the stored fit: 1 at rungs = 1)
    if beta is None:
        beta = dg["rung_details"]["thermodynamic_power"].to_numpy(dtype=np.float64)
    parameters = {
        "data": r_list({k: r_real(v) for k, v in pr["data"].items()}),
        "df_params": r_dataframe(pr["df_params"].astype({"trans_type": np.float64}), ("tbl_df", "tbl", "data.frame")),
        "loglike": r_str(pr["loglike"]), "logprior": r_str(pr["logprior"]), "burnin": r_real(float(pr["burnin"])),
        "samples": r_real(float(pr["samples"])), "rungs": r_int(pr["rungs"]), "chains": r_real(float(pr["chains"])),
        "coupling_on": r_lgl(bool(pr["coupling_on"])), "GTI_pow": r_real(float(pr["GTI_pow"])),
        "beta_manual": r_real(beta),
    }
    mcmcout = r_list({"output": r_dataframe(mc["output"]), "diagnostics": r_list(diagnostics), "parameters": r_list(parameters)},
                     cls="drjacoby_output")
    return r_list({"inputs": r_list(inputs), "mcmcout": mcmcout}, cls="IFRmodel_inf")


def write_fit_rds(path, fit, compress=True):
    """saveRDS() equivalent for the result of `api.run_IFRmodel_age`."""
    write_rds(path, fit_to_r(fit), compress=compress)


def r_to_python(o):
    """Inverse of `to_r` for the types the fit object uses (tests, and loading a fit back on a box without R)."""
    import pandas as pd
    if o is None:
        return None
    t, attr = o["t"], o.get("attr", {})
    if t == 19:
        nm = attr.get("names", {}).get("v")
        cls = attr.get("class", {}).get("v", [])
        if "data.frame" in cls:
            return pd.DataFrame({n: _column(c) for n, c in zip(nm, o["v"])})
        vals = [r_to_python(v) for v in o["v"]]
        return dict(zip(nm, vals)) if nm else vals
    v = _column(o)
    if "dim" in attr:
        return np.asarray(v).reshape(attr["dim"]["v"], order="F")
    if "names" in attr and t == 14:
        return pd.Series(v, index=attr["names"]["v"])
    return v[0] if len(v) == 1 else v


def _column(o):
    t, v = o["t"], o["v"]
    if t == 16:
        return np.array(v, dtype=object)
    if t == 14:
        return np.asarray(v, dtype=np.float64)
    if t == 13:
        if "levels" in o.get("attr", {}):
            lv = o["attr"]["levels"]["v"]
            return np.array([None if k == NA_INTEGER else lv[k - 1] for k in v], dtype=object)
        return np.asarray(v, dtype=np.int64)
    if t == 10:
        return np.array([None if k == NA_INTEGER else bool(k) for k in v], dtype=object)
    raise ValueError("unsupported column type %d" % t)
