"""Command line mirroring experiments/experiments.py:21-34 (`fire` is replaced by a tiny parser):
    python -m mopoe_b200.cli train --dataset hbn --datasetdir D --outdir O --input_dims 7,444 ...
    python -m mopoe_b200.cli daa   --dataset hbn --datasetdir D --outdir O --run hbn_2026_... ...
    python -m mopoe_b200.cli rsa   --dataset hbn --datasetdir D --outdir O --run hbn_2026_... ...
Every `--key value` becomes a keyword argument of workflow.train_exp / daa_exp / rsa_exp."""
import ast
import sys


def _parse(argv):
    kw, i = {}, 0
    while i < len(argv):
        key = argv[i].lstrip("-").replace("-", "_")
        if "=" in key:
            key, val = key.split("=", 1)
        else:
            i += 1
            val = argv[i] if i < len(argv) else "True"
        try:
            kw[key] = ast.literal_eval(val)
        except Exception:
            kw[key] = val
        i += 1
    return kw


def main(argv=None):
    from . import workflow
    argv = list(sys.argv[1:] if argv is None else argv)
    commands = {"train": workflow.train_exp, "daa": workflow.daa_exp, "rsa": workflow.rsa_exp}
    if not argv or argv[0] not in commands:
        raise SystemExit("usage: cli.py {train|daa|rsa} --key value ...   (the other reference commands are out of scope)")
    kw = _parse(argv[1:])
    if isinstance(kw.get("input_dims"), tuple):
        kw["input_dims"] = list(kw["input_dims"])
    return commands[argv[0]](**kw)


if __name__ == "__main__":
    main()
