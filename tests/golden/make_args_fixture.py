"""Fixture for the build(args) drop-in test: the reference's own argument parser and the state_dict it builds.

    python tests/golden/make_args_fixture.py      (build container only: needs /root/reference)

get_args_parser is exec-extracted from /root/reference/main.py:35-147 (the module itself imports pycocotools);
for each argument set the UNMODIFIED reference build(args) (through oracle/ref_shim.py) gives the state_dict key -> shape map.
Writes tests/golden/build_args.json: {"defaults": {...}, "cases": {name: {"argv": [...], "args": {parsed}, "state": {key: shape}}}}.
"""
import argparse
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# the model-relevant arguments of scripts/run_coco17.py:12-36 and scripts/run_voc0712.py (data paths / schedule left out)
CASES = {
    "parser_defaults": [],
    "run_coco17": ["--dataset_file", "coco", "--fixed_size", "--enc_layers", "3", "--layer_to_det", "24", "--focal_gamma", "0.5",
                   "--backbone", "TSCAM_cait_XXS36_Two_Branch", "--max_size", "512", "--num_queries", "300", "--backbone_drop_rate", "0.07",
                   "--drop_path_rate", "0.2", "--drop_attn_rate", "0.05", "--hungarian_multi", "--hung_match_ratio", "5"],
}


def parser_from_reference():
    src = open("/root/reference/main.py").read()
    m = re.search(r"^def get_args_parser\(\):.*?^    return parser\n", src, re.S | re.M)
    ns = {"argparse": argparse}
    exec(m.group(0), ns)
    return ns["get_args_parser"]()


def main():
    assert ref_shim.available(), "needs /root/reference"
    parser = parser_from_reference()
    defaults = {k: v for k, v in vars(parser.parse_args([])).items() if isinstance(v, (int, float, str, bool, type(None), list))}
    ref = ref_shim.load_reference()
    out = {"defaults": defaults, "cases": {}}
    for name, argv in CASES.items():
        args = parser.parse_args(argv)
        args.device = "cpu"
        torch.manual_seed(0)
        model = ref.conditional_detr.build(args)[0]
        parsed = {k: v for k, v in vars(args).items() if isinstance(v, (int, float, str, bool, type(None), list))}
        out["cases"][name] = {"argv": argv, "args": parsed, "state": {k: list(v.shape) for k, v in model.state_dict().items()}}
        print(name, len(out["cases"][name]["state"]), "keys")
    with open(os.path.join(HERE, "build_args.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
