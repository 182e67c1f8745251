"""Genotype decoder + derived-network builder against fuzz goldens produced by the UNMODIFIED reference
(oracle/make_golden_decode.py): 120 random architectures x (3 decodes + 8 `lasts` choices).  Checked for the product class
(`fasterseg_b200.model_seg.Network_Multi_Path_Infer`) and for the oracle's own restatement of the decoder."""
import pytest

from oracle import fasterseg_oracle as orc
from oracle import make_golden_decode as mk
from tests import helpers as H

FUZZ = H.load_json("decode_fuzz.json")
SEEDS = sorted(FUZZ["cases"], key=int)


@pytest.mark.parametrize("chunk", range(6))
def test_product_decoder_and_builder_match_reference_fuzz(chunk, monkeypatch):
    from fasterseg_b200 import operations
    from fasterseg_b200.model_seg import Network_Multi_Path_Infer
    monkeypatch.setattr(operations, "latency_lookup_table", mk.SyntheticLatencyTable())
    for seed in SEEDS[chunk::6]:
        entry = FUZZ["cases"][seed]
        got = mk.run_case(Network_Multi_Path_Infer, mk.draw_case(int(seed)), entry["training"])
        want = entry["rec"]
        assert got.keys() == want.keys(), seed
        for last in ("0", "1", "2"):
            g, w = got["decoded"][last], want["decoded"][last]
            assert g["ops"] == w["ops"] and g["path"] == w["path"] and g["downs"] == w["downs"], (seed, last)
            assert g["widths"] == pytest.approx(w["widths"], abs=1e-12), (seed, last)
        assert got["alphas_neg_inf"] == want["alphas_neg_inf"], seed
        for lasts, w in want["structures"].items():
            g = dict(got["structures"][lasts])
            w = dict(w)
            # forward_latency over the synthetic table: same keys looked up, same summation order
            assert g.pop("latency_1024x2048") == pytest.approx(w.pop("latency_1024x2048"), rel=1e-12, abs=1e-12), (seed, lasts)
            assert g == w, (seed, lasts)


def test_oracle_decoder_matches_reference_fuzz():
    for seed in SEEDS:
        case = mk.draw_case(int(seed))
        want = FUZZ["cases"][seed]["rec"]
        alphas, betas, ratios = mk.clone_params(case)
        wml = mk.WML
        if ratios[0].size(1) == 1:
            wml = [1.] if case["ignore_skip"] else [4. / 12]
        for last in (0, 1, 2):
            ops, path, downs, widths = orc.network_metas(alphas, betas, ratios, wml, case["layers"], last, case["ignore_skip"])
            w = want["decoded"][str(last)]
            assert [int(o) for o in ops] == w["ops"] and list(path) == w["path"] and list(downs) == w["downs"], (seed, last)
            assert [float(x) for x in widths] == pytest.approx(w["widths"], abs=1e-12), (seed, last)
