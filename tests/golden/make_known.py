"""Extract the reference's known-answer checks into JSON fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_known.py
Source: /root/reference/test_suite/tests/tests_known.rs (one test, 44 duration checks, state carried from one
check to the next). Output: tests/golden/known_trajectories.json — an ordered op list that tests/test_known.py replays.
Only numbers and the order of operations are extracted; no reference code is copied.
"""
import json
import os
import re
import sys

SRC = "/root/reference/test_suite/tests/tests_known.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "known_trajectories.json")


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def ev(expr):
    e = expr.strip().replace("std::f64::consts::FRAC_PI_2", "1.5707963267948966").replace("std::f64::consts::PI", "3.141592653589793")
    e = e.replace("f64::INFINITY", "float('inf')").replace("true", "True").replace("false", "False")
    return eval(e, {"__builtins__": {}}, {"float": float})


def ev_list(body):
    parts = [p for p in (x.strip() for x in body.split(",")) if p]
    return [ev(p) for p in parts]


def main():
    text = strip_comments(open(SRC).read())
    body = text[text.index("fn test_known_trajectory()"):]
    body = body[body.index("{") + 1:]
    body = re.sub(r"\[([^\[\];]+); (\d+)\]", lambda m: "[" + ", ".join([m.group(1).strip()] * int(m.group(2))) + "]", body)
    stmts = [re.sub(r"\s+", " ", s).strip() for s in body.split(";")]
    ops = []
    unknown = []
    for s in stmts:
        if not s or s == "}":
            continue
        m = re.match(r"let mut (\w+) = Ruckig::<(\d+), (\w+)>::new\(None, ([^)]+)\)$", s)
        if m:
            ops.append({"op": "new_otg", "var": m.group(1), "dof": int(m.group(2)), "handler": m.group(3), "delta_time": ev(m.group(4))})
            continue
        m = re.match(r"let mut (\w+) = InputParameter::new\(None\)$", s)
        if m:
            ops.append({"op": "new_input", "var": m.group(1)})
            continue
        m = re.match(r"(\w+)\.(\w+) = (Some\()?DataArrayOrVec::Stack\(\[(.*)\]\)\)?$", s)
        if m:
            ops.append({"op": "set", "var": m.group(1), "field": m.group(2), "value": ev_list(m.group(4))})
            continue
        m = re.match(r"(\w+)\.(\w+) = None$", s)
        if m:
            ops.append({"op": "set", "var": m.group(1), "field": m.group(2), "value": None})
            continue
        m = re.match(r"(\w+)\.(\w+) = Some\(([^()]+)\)$", s)
        if m:
            ops.append({"op": "set", "var": m.group(1), "field": m.group(2), "value": ev(m.group(3))})
            continue
        m = re.match(r"(\w+)\.(\w+) = (\w+)::(\w+)$", s)
        if m:
            ops.append({"op": "set_enum", "var": m.group(1), "field": m.group(2), "enum": m.group(3), "value": m.group(4)})
            continue
        m = re.match(r"check_(full_)?duration\(&mut (\w+), &(mut )?(\w+), ([^)]+)\)$", s)
        if m:
            ops.append({"op": "check_full_duration" if m.group(1) else "check_duration", "otg": m.group(2), "input": m.group(4), "duration": ev(m.group(5))})
            continue
        m = re.match(r"let mut (\w+) = Trajectory::<(\d+)>::new\(None\)$", s)
        if m:
            ops.append({"op": "new_trajectory", "var": m.group(1), "dof": int(m.group(2))})
            continue
        m = re.match(r"(\w+)\.calculate\(&(\w+), &mut (\w+)\)\.unwrap\(\)$", s)
        if m:
            ops.append({"op": "calculate", "otg": m.group(1), "input": m.group(2), "traj": m.group(3)})
            continue
        m = re.match(r"(\w+)\.at_time\( (\w+)\.get_duration\(\) / 2\.0,", s)
        if m:
            ops.append({"op": "at_half_duration", "traj": m.group(1)})
            continue
        m = re.match(r"assert_float_eq!\(new_position\[(\d+)\], ([^,]+), abs <= ([^)]+)\)$", s)
        if m:
            ops.append({"op": "assert_new_position", "index": int(m.group(1)), "value": ev(m.group(2)), "abs": ev(m.group(3))})
            continue
        if re.match(r"let mut new_(position|velocity|acceleration) = ", s):
            continue
        unknown.append(s)
    if unknown:
        print("UNPARSED STATEMENTS:", file=sys.stderr)
        for u in unknown:
            print("   ", u[:200], file=sys.stderr)
        sys.exit(1)
    n_checks = sum(1 for o in ops if o["op"].startswith("check_"))
    json.dump({"source": "rsruckig v2.1.2 test_suite/tests/tests_known.rs", "tolerance_abs": 1e-4, "checks": n_checks, "ops": ops},
              open(OUT, "w"), indent=0)
    print(f"{len(ops)} ops, {n_checks} duration checks -> {OUT}")


if __name__ == "__main__":
    main()
