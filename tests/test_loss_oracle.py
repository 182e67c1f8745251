"""Oracle of the callers' loss path ("next" row N1) against goldens from the UNMODIFIED reference
(oracle/make_golden_loss.py): ProbOhemCrossEntropy2d over its decision branches and the KL distillation term, loss value
and gradient w.r.t. the logits."""
import math

import numpy as np
import pytest
import torch

from oracle import fasterseg_oracle as orc
from oracle import make_golden_loss as mk
from tests import helpers as H

Z = H.load_npz("loss.npz")


@pytest.mark.parametrize("name", sorted(mk.OHEM_CASES))
def test_ohem_cross_entropy_matches_reference(name):
    pred, tgt, thresh, min_kept = mk.ohem_inputs(name)
    pred.requires_grad_(True)
    loss = orc.ohem_cross_entropy(pred, tgt, ignore_label=255, thresh=thresh, min_kept=min_kept)
    want = float(Z[name + "/loss"][0])
    if math.isnan(want):          # every label ignored: the reference's CrossEntropyLoss returns nan, so must the restatement
        assert math.isnan(float(loss))
        return
    assert float(loss) == pytest.approx(want, rel=1e-6)
    loss.backward()
    np.testing.assert_allclose(pred.grad.numpy()[:, :, ::4, ::4], Z[name + "/grad.s4"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", sorted(mk.KL_CASES))
def test_distillation_kl_matches_reference(name):
    student, teacher = mk.kl_inputs(name)
    student.requires_grad_(True)
    loss = orc.distill_kl(student, teacher)
    assert float(loss) == pytest.approx(float(Z[name + "/loss"][0]), rel=1e-5)
    loss.backward()
    np.testing.assert_allclose(student.grad.numpy()[:, :, ::4, ::4], Z[name + "/grad.s4"], rtol=1e-4, atol=1e-10)


# ---- the product's own criteria (fasterseg_b200/losses.py, sync-free restatement) against the same reference goldens ----------
@pytest.mark.parametrize("name", sorted(mk.OHEM_CASES))
def test_product_ohem_matches_reference(name):
    from fasterseg_b200.losses import ProbOhemCrossEntropy2d
    pred, tgt, thresh, min_kept = mk.ohem_inputs(name)
    pred.requires_grad_(True)
    loss = ProbOhemCrossEntropy2d(ignore_label=255, thresh=thresh, min_kept=min_kept)(pred, tgt)
    want = float(Z[name + "/loss"][0])
    if math.isnan(want):
        assert math.isnan(float(loss))
        return
    assert float(loss) == pytest.approx(want, rel=1e-6)
    loss.backward()
    np.testing.assert_allclose(pred.grad.numpy()[:, :, ::4, ::4], Z[name + "/grad.s4"], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("name", sorted(mk.KL_CASES))
def test_product_distillation_kl_matches_reference(name):
    from fasterseg_b200.losses import distillation_kl
    student, teacher = mk.kl_inputs(name)
    student.requires_grad_(True)
    loss = distillation_kl(student, teacher)
    assert float(loss) == pytest.approx(float(Z[name + "/loss"][0]), rel=1e-5)
    loss.backward()
    np.testing.assert_allclose(student.grad.numpy()[:, :, ::4, ::4], Z[name + "/grad.s4"], rtol=1e-4, atol=1e-10)
