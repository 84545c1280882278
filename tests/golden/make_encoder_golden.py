"""Generates tests/golden/encoder_golden.npz by RUNNING THE REFERENCE'S OWN ENCODER in the build container.

    python tests/golden/make_encoder_golden.py          (needs /root/reference; CPU only)

The reference module neuroestimator/estimator/encoder.py imports only pandas / numpy / networkx, so -- unlike
the NNGP arithmetic -- it can be imported here.  A synthetic 3-table schema (numerical + categorical columns,
two PK/FK join keys) is built, random query lines in the wire format of neuroestimator/README.md:35-48 are
encoded with NNGPEncoder.parse_line_without_card_then_encode (encoder.py:229-250) and NNGPEncoder.parse_line +
transform_to_1d_array (encoder.py:207-227, 197-205), and the lines + the resulting float64 rows are stored.
The single-table format of Queries/forest_data is covered by GeneralQuerySampler-equivalent rows taken from
tests/golden/forest_xy.npz (see make_forest_fixture.py).
"""
import contextlib
import importlib.util
import io
import sys
from pathlib import Path

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "nngp-src_b200"))
REF = Path("/root/reference/neuroestimator/estimator/encoder.py")


def load_reference_encoder():
    spec = importlib.util.spec_from_file_location("ref_encoder", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_encoder()
    rng = np.random.default_rng(7)
    n = 400
    orders = pd.DataFrame({"o_key": np.arange(n), "c_key": rng.integers(0, 50, n), "o_total": rng.uniform(5, 9000, n),
                           "o_prio": rng.integers(0, 7, n), "o_date": rng.integers(19920101, 19981231, n)})
    cust = pd.DataFrame({"c_key": np.arange(50), "c_bal": rng.uniform(-999, 9999, 50), "c_seg": rng.integers(0, 140, 50)})
    items = pd.DataFrame({"o_key": rng.integers(0, n, 900), "l_qty": rng.integers(1, 51, 900).astype(float),
                          "l_price": rng.uniform(900, 105000, 900), "l_const": np.full(900, 3.0)})
    types = {"orders": ["categorical", "categorical", "numerical", "categorical", "numerical"],
             "customer": ["categorical", "numerical", "categorical"],
             "lineitem": ["categorical", "numerical", "numerical", "numerical"]}
    with contextlib.redirect_stdout(io.StringIO()):
        tables = [ref.Table(orders, types["orders"], "orders", chunk_size=64),
                  ref.Table(cust, types["customer"], "customer", chunk_size=64),
                  ref.Table(items, types["lineitem"], "lineitem", chunk_size=64)]
        enc = ref.NNGPEncoder(tables)
    from nngp_b200.encoder import schema_text_from_reference_encoder
    schema = schema_text_from_reference_encoder(enc)

    def num_pred(df, col):
        a, b = sorted(rng.uniform(df[col].min(), df[col].max(), 2))
        style = rng.integers(0, 3)
        if style == 0:
            return f"{col},{b:.3f},{a:.3f}"
        if style == 1:
            return f"{col}, {int(b)} ,{int(a)}"
        return f"{col},{float(b)!r},{float(a)!r}"

    def cat_pred(tab, col):
        ncat = len(tab.categorical_codes_dict[col])
        k = int(rng.integers(1, min(ncat, 6) + 1))
        cats = rng.choice(ncat, size=k, replace=False)
        return col + "," + ",".join(str(int(c)) for c in cats)

    tabs = {"orders": (tables[0], orders), "customer": (tables[1], cust), "lineitem": (tables[2], items)}
    joins = {("orders", "customer"): "c_key", ("orders", "lineitem"): "o_key"}
    lines = []
    for _ in range(600):
        k = int(rng.integers(1, 4))
        if k == 1:
            names = [str(rng.choice(list(tabs)))]
        elif k == 2:
            names = list(rng.choice([["orders", "customer"], ["lineitem", "orders"], ["customer", "orders"]]))
        else:
            names = list(rng.permutation(["orders", "customer", "lineitem"]))
        terms = [",".join(names)]
        for nm in names:
            tab, df = tabs[nm]
            preds = []
            for i, col in enumerate(df.columns):
                if rng.random() < 0.45:
                    preds.append(cat_pred(tab, col) if tab.col_types[i] == "categorical" else num_pred(df, col))
            terms.append("#".join(preds))
        js = []
        for (a, b), col in joins.items():
            if a in names and b in names and rng.random() < 0.9:
                pair = (a, b) if rng.random() < 0.5 else (b, a)
                js.append(f"{pair[0]},{pair[1]},{col}" + (",=" if rng.random() < 0.3 else ""))
        terms.append("#".join(js))
        lines.append("@".join(terms))
    x = np.array([enc.parse_line_without_card_then_encode(l) for l in lines])
    cards = rng.integers(1, 10**7, len(lines))
    train_lines = [f"{l}@{c}" for l, c in zip(lines, cards)]
    xt = []
    for l in train_lines[:100]:
        table_ids, all_pred_list, join_infos, card = enc.parse_line(l)
        xt.append(enc.transform_to_1d_array(table_ids, all_pred_list, join_infos))
    # --- single-table format (Queries/forest_data) through the reference's GeneralQuerySampler ------------------
    # QuerySampler.py imports `datasets` and `util` (-> pandasql / clickhouse / seaborn, absent here) only for its
    # data-generation paths; they are stubbed so that the class's own parse_line (:157-170) and
    # transform_to_1d_array (:200-221) run unmodified.  forest.csv is not shipped (readme.md:37): a two-row frame
    # carrying the UCI Covertype min/max per column gives the same all_col_ranges (QuerySampler.py:50-51).
    import types
    sys.modules.setdefault("datasets", types.ModuleType("datasets"))
    util_stub = types.ModuleType("util")
    util_stub.make_dir = lambda *_a, **_k: None
    sys.modules.setdefault("util", util_stub)
    spec = importlib.util.spec_from_file_location("ref_query_sampler", "/root/reference/QuerySampler.py")
    qmod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(qmod)
    ranges = {"A": (1859, 3858), "B": (0, 360), "C": (0, 66), "D": (0, 1397), "E": (-173, 601),
              "F": (0, 7117), "G": (0, 254), "H": (0, 254), "I": (0, 254), "J": (0, 7173)}
    fdf = pd.DataFrame({k: [float(v[0]), float(v[1])] for k, v in ranges.items()})
    with contextlib.redirect_stdout(io.StringIO()):
        qs = qmod.GeneralQuerySampler(fdf, ["numerical"] * 10, "forest", chunk_size=64)
    from nngp_b200.encoder import schema_text_from_query_sampler
    forest_schema = schema_text_from_query_sampler(qs)
    forest_lines = []
    for name in ("query_2.txt", "query_5.txt", "query_10.txt"):
        with open(f"/root/reference/Queries/forest_data/{name}") as fh:
            forest_lines += [l.strip() for l in fh.readlines()[:120]]
    fx, fcard = [], []
    for l in forest_lines:
        pred_list, card = qs.parse_line(l)
        fx.append(qs.transform_to_1d_array(pred_list))
        fcard.append(card)
    np.savez_compressed(Path(__file__).with_name("encoder_golden.npz"), schema=np.array(schema), lines=np.array(lines),
                        x=x, train_lines=np.array(train_lines[:100]), x_train=np.array(xt), cards=cards[:100].astype(np.float64),
                        forest_schema=np.array(forest_schema), forest_lines=np.array(forest_lines), forest_x=np.array(fx),
                        forest_cards=np.array(fcard, dtype=np.float64))
    print("schema:\n" + schema)
    print("dim", x.shape, "nonzero frac", float((x != 0).mean()), "max", x.max())


if __name__ == "__main__":
    main()
