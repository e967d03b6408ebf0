"""Regenerates tests/golden/stage1_fixtures.json from the reference's own stage-1 fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Fixture format (reference tests/test_stage_1.mojo:85-96): line 1 = JSON text, line 2 = a mask
string with '1' under every expected structural character.  We store the input text, the index
list derived from line 2, and the trailer the reference's test asserts (:70-82).
"""
import json
import os

REF = "/root/reference/tests/jsons_for_test"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage1_fixtures.json")


def load(path, harness_negative):
    with open(path, "r", encoding="utf-8") as f:
        lines = f.read().splitlines()
    text, mask = lines[0], lines[1]
    claimed = [i for i, c in enumerate(mask) if c == "1"]
    return {
        "file": os.path.relpath(path, REF),
        "input": text,
        "mask": mask,
        "claimed_indexes": claimed,
        "len": len(text.encode("utf-8")),
        "expect_error": 0,
        # the two top-level files are the reference harness's self-tests: their line 2 is
        # deliberately wrong (test_stage_1.mojo:99-110), so the claimed indexes must NOT match
        "harness_negative": harness_negative,
    }


def main():
    out = []
    vdir = os.path.join(REF, "valid")
    for name in sorted(os.listdir(vdir)):
        out.append(load(os.path.join(vdir, name), False))
    for name in ("wrong_tagging.json", "detect_incorrect_result.json"):
        out.append(load(os.path.join(REF, name), True))
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump({"source": "gabrieldemarmiesse/mojo-simdjson tests/jsons_for_test", "fixtures": out}, f, indent=1)
    print(f"wrote {len(out)} fixtures to {OUT}")


if __name__ == "__main__":
    main()
