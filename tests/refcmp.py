"""helpers: compare an alignment result (api.unpack_result dict) with an oracle dump (oracle.runner.parse_dump)"""
import numpy as np


def result_to_dump(res):
    mums = []
    for i in range(len(res["mum_length"])):
        mums.append((int(res["mum_length"][i]), int(res["mum_slength"][i]),
                     [(int(s), int(e), int(f)) for s, e, f in zip(res["mum_start"][i], res["mum_end"][i], res["mum_fwd"][i])]))
    cl = []
    for i in range(len(res["cluster_type"])):
        cl.append((int(res["cluster_type"][i]), int(res["cluster_nmums"][i]), int(res["cluster_length"][i]),
                   [(int(s), int(e)) for s, e in zip(res["cluster_start"][i], res["cluster_end"][i])]))
    return dict(n=res["n"], mums=mums, clusters=cl)


def write_dump(res, path):
    """a result of the product in the format of the reference's hook dump (oracle/build_ref.py H1) plus one "I" line per
    cluster: the indices of its MUMs in the list - input of oracle/_ref/xmfa_from_dump"""
    d = result_to_dump(res)
    off, idx = res["cluster_mum_off"], res["cluster_mum_idx"]
    with open(path, "w") as f:
        f.write("N %d\n" % d["n"])
        for ln, sl, cols in d["mums"]:
            f.write("M %d %d %s\n" % (ln, sl, " ".join("%d:%d:%d" % c for c in cols)))
        for k, (t, nm, ln, cols) in enumerate(d["clusters"]):
            f.write("C %d %d %d %s\n" % (t, nm, ln, " ".join("%d:%d" % c for c in cols)))
            f.write("I %s\n" % " ".join(str(int(x)) for x in idx[off[k]:off[k + 1]]))


def diff_dumps(a, b, limit=5):
    """returns list of human-readable differences (empty = identical)"""
    out = []
    if a["n"] != b["n"]:
        out.append("n: %s vs %s" % (a["n"], b["n"]))
    if len(a["mums"]) != len(b["mums"]):
        out.append("#mums: %d vs %d" % (len(a["mums"]), len(b["mums"])))
    for i, (x, y) in enumerate(zip(a["mums"], b["mums"])):
        if x != y:
            out.append("mum %d: %s vs %s" % (i, x, y))
            if len(out) > limit:
                break
    if len(a["clusters"]) != len(b["clusters"]):
        out.append("#clusters: %d vs %d" % (len(a["clusters"]), len(b["clusters"])))
    for i, (x, y) in enumerate(zip(a["clusters"], b["clusters"])):
        if x != y:
            out.append("cluster %d: %s vs %s" % (i, x, y))
            if len(out) > 2 * limit:
                break
    return out


def dump_lines(res):
    """the product's result as the lines of the reference's hook dump (oracle/build_ref.py H1: 'N', 'M ...', 'C ...'), without
    newlines - vectorised, for full-size results (10^5..10^6 MUMs)"""
    yield "N %d" % res["n"]
    ln, sl = res["mum_length"], res["mum_slength"]
    st, en, fw = res["mum_start"], res["mum_end"], res["mum_fwd"]
    n = st.shape[1] if len(ln) else 0
    for i in range(len(ln)):
        yield "M %d %d " % (ln[i], sl[i]) + " ".join("%d:%d:%d" % (st[i, k], en[i, k], fw[i, k]) for k in range(n))
    ct, cn, cl = res["cluster_type"], res["cluster_nmums"], res["cluster_length"]
    cs, ce = res["cluster_start"], res["cluster_end"]
    for i in range(len(ct)):
        yield "C %d %d %d " % (ct[i], cn[i], cl[i]) + " ".join("%d:%d" % (cs[i, k], ce[i, k]) for k in range(cs.shape[1]))
