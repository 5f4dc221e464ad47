"""Drop-in boundary: `from spe_b200.models import build_model; build_model(args)` with the reference's OWN arguments -- the parser
defaults (main.py:35-147) and the training script's arguments (scripts/run_coco17.py:12-36, incl. its non-zero drop rates) --
builds the state_dict the unmodified reference builds (tests/golden/build_args.json, written by make_args_fixture.py).  CPU only."""
import argparse
import json
import os

import pytest
import torch


def args_from_fixture(fix, case, device="cpu"):
    """the Namespace the reference's parser produced for this case (parsed in the build container, stored in the fixture)"""
    d = dict(fix["cases"][case]["args"])
    d["device"] = device
    return argparse.Namespace(**d)


@pytest.fixture(scope="module")
def fixture(golden_dir):
    return json.load(open(os.path.join(golden_dir, "build_args.json")))


@pytest.mark.parametrize("case", ["parser_defaults", "run_coco17"])
def test_build_model_with_reference_args(fixture, case):
    from spe_b200.models import build_model
    c = fixture["cases"][case]
    args = args_from_fixture(fixture, case)
    if case == "run_coco17":
        assert args.backbone_drop_rate == 0.07 and args.drop_path_rate == 0.2 and args.drop_attn_rate == 0.05 and args.dropout == 0.1
    torch.manual_seed(0)
    model, criterion, criterion_refine, post, post_refine = build_model(args)
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert set(mine) == set(c["state"]), (sorted(set(mine) ^ set(c["state"]))[:10])
    bad = [(k, mine[k], c["state"][k]) for k in mine if mine[k] != c["state"][k]]
    assert not bad, bad[:5]
    names = [n for n, _ in model.named_parameters()]
    assert any("backbone" in n for n in names) and any("blocks_token_only" in n for n in names)     # optimizer groups, main.py:177-186
    assert criterion.losses == ["labels", "boxes", "cardinality", "image_label"] and criterion_refine.losses == ["labels", "boxes", "cardinality"]
    assert set(post) == {"bbox"} and set(post_refine) == {"bbox"}
    assert criterion.matcher.match_ratio == (args.hung_match_ratio if args.hungarian_multi else 1) or hasattr(criterion.matcher, "match_ratio")
