"""Pin the RESULT SCHEMA of the reference drivers (they cannot be imported or run: module-level argparse and the
bugs of SURVEY 3.7) by reading their source with `ast`: the keys of every dict the drivers append to their result
lists, the pickle file-name fragments and the class-score file-name fragments.  Runs where /root/reference exists;
writes tests/golden/result_schema.json, which the CPU tests compare the drop-in drivers' output against.
    python oracle/pin_result_schema.py"""
import ast
import json
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/video_features_pytorch"


def extract(path):
    tree = ast.parse(open(path).read())
    appended, strings = {}, []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "append" \
                and node.args and isinstance(node.args[0], ast.Dict):
            target = ast.unparse(node.func.value)
            appended[target] = [k.value for k in node.args[0].keys]
        if isinstance(node, ast.Constant) and isinstance(node.value, str):
            s = node.value
            if any(t in s for t in ("Results", "ClassScore", "cam_saved_images", "combined", "_gs%", "_cs%", "g_")) \
                    and len(s) < 60:
                strings.append(s)
    return {"appended_dict_keys": appended, "name_fragments": sorted(set(strings))}


def main():
    out = {"smth": extract(os.path.join(REF, "FindMasksComparison_I3D_smth.py")),
           "kth": extract(os.path.join(REF, "FindMasksComparison_I3D_KTH.py"))}
    # the command-line flags of pt/utils.py:12-91 (long names), for the drop-in parser
    tree = ast.parse(open(os.path.join(REF, "utils.py")).read())
    flags = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr == "add_argument":
            flags.append([a.value for a in node.args if isinstance(a, ast.Constant)])
    out["cli_flags"] = flags
    # config keys of the shipped config files
    cfg = {}
    for name in ("config_i3d_smth.py", "config_i3d_kth.py", "config_clstm_kth.py"):
        ns = {}
        exec(open(os.path.join(REF, "configs", name)).read(), ns)
        cfg[name] = sorted(ns["config"].keys())
    out["config_keys"] = cfg
    path = os.path.join(REPO, "tests", "golden", "result_schema.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)
    print(json.dumps(out["smth"], indent=1)[:1500])


if __name__ == "__main__":
    main()
