"""Child process of tests/test_gpu_corrupt.py: decodes one (possibly damaged) file with the library and reports how it ended.
A damaged stream may decode to garbage or be refused, but the call must RETURN (the parent enforces a timeout)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from fuif_b200 import api  # noqa: E402


def main():
    path, index_json = sys.argv[1], sys.argv[2]
    data = open(path, "rb").read()
    index = json.loads(index_json) if index_json != "-" else None
    if index is not None:
        index = (index[0], index[1])
    ctx = api.Context(0)
    out = {"decoded": False, "undone": False, "error": None}
    try:
        img = api.fuif_decode(data, ctx=ctx, group_index=index)
        out["decoded"] = True
        img.undo_transforms(0)
        ctx.synchronize()
        out["undone"] = True
    except api.FuifError as e:
        out["error"] = str(e)[:200]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
