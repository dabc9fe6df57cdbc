"""Parity report: every -m gpu parity test records its worst error ratios here; the session writes them to
gpurun_out/parity_report.json (scratch; the copy to keep is committed under profiles/)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_rows = []


def record(test, **vals):
    _rows.append(dict(test=test, **vals))


def dump():
    if not _rows:
        return
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_report.json")
        old = []
        if os.path.exists(path):
            try:
                with open(path) as f:
                    old = json.load(f)
            except Exception:
                old = []
        seen = {r["test"] for r in _rows}
        with open(path, "w") as f:
            json.dump([r for r in old if r.get("test") not in seen] + _rows, f, indent=1)
    except OSError:
        pass
