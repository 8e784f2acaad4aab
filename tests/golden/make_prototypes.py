"""Writes tests/golden/reference_prototypes.json: the normalised prototypes of every function the reference's public headers
declare (include/fft_auto.h, include/fft_gpu.h under /root/reference), plus the values of its public enums and flag macros.
tests/test_abi.py compares include/*.h of this repo against it declaration by declaration (the reference tree does not
travel to the GPU box). Interface facts only - no code. usage: python tests/golden/make_prototypes.py"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import proto_parse  # noqa: E402

REF = "/root/reference/include"


def main():
    res = {"prototypes": {}, "constants": {}}
    for h in ("fft_auto.h", "fft_gpu.h", "fft_common.h"):
        text = open(os.path.join(REF, h)).read()
        if h != "fft_common.h":
            res["prototypes"][h] = proto_parse.prototypes(text)
        res["constants"][h] = proto_parse.constants(text)
    with open(os.path.join(HERE, "reference_prototypes.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in res["prototypes"].items()}, {k: len(v) for k, v in res["constants"].items()})


if __name__ == "__main__":
    main()
