#!/usr/bin/env python
"""Diff kubeflow_b200.suggestion.api_pb.SCHEMA against a real katib api.proto.

    python tools/check_proto.py path/to/kubeflow/katib/pkg/apis/manager/v1beta1/api.proto     # exit 1 on any mismatch
    python tools/check_proto.py --emit > /tmp/schema.proto                                      # SCHEMA rendered as .proto text

The field numbers in api_pb.py were written down from memory of upstream (the proto is not in /root/reference and cannot be
fetched); a wrong number drops a field silently against a real katib-controller.  This tool parses the message / enum
definitions of a proto3 file (nested messages, map<string,string>, repeated, comments) and reports, for every message the
suggestion surface uses, fields whose number, type, label or name differ, fields missing on either side, and enum value
differences.  No protoc needed."""
from __future__ import annotations

import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kubeflow_b200.suggestion import api_pb  # noqa: E402

T = api_pb.T
TYPE_NAMES = {T.TYPE_STRING: "string", T.TYPE_INT32: "int32", T.TYPE_DOUBLE: "double"}


def schema_as_dict():
    """{message: {field name: (number, type string, repeated?)}}, {enum: [names]} from api_pb."""
    msgs = {}
    for full, fields in api_pb.SCHEMA.items():
        d = {}
        for name, num, typ, label, tname in fields:
            if typ == "map":
                ts = "map<string,string>"
            elif typ in (T.TYPE_MESSAGE, T.TYPE_ENUM):
                ts = tname.split(".")[-1]
            else:
                ts = TYPE_NAMES[typ]
            d[name] = (num, ts, label == T.LABEL_REPEATED and typ != "map")
        msgs[full] = d
    enums = dict(api_pb.ENUMS)
    enums["TrialStatus.TrialConditionType"] = list(api_pb.TRIAL_CONDITIONS)
    return msgs, enums


def emit():
    msgs, enums = schema_as_dict()
    out = ['syntax = "proto3";', f"package {api_pb.PACKAGE};", ""]
    for e, vals in enums.items():
        if "." in e:
            continue
        out += [f"enum {e} {{"] + [f"  {v} = {i};" for i, v in enumerate(vals)] + ["}", ""]

    def render(full, indent):
        pad = "  " * indent
        lines = [f"{pad}message {full.split('.')[-1]} {{"]
        for child in [m for m in msgs if m.startswith(full + ".") and m.count(".") == full.count(".") + 1]:
            lines += render(child, indent + 1)
        for e, vals in enums.items():
            if e.startswith(full + "."):
                lines += [f"{pad}  enum {e.split('.')[-1]} {{"] + [f"{pad}    {v} = {i};" for i, v in enumerate(vals)] + [f"{pad}  }}"]
        for name, (num, ts, rep) in msgs[full].items():
            lines.append(f"{pad}  {'repeated ' if rep else ''}{ts} {name} = {num};")
        lines.append(f"{pad}}}")
        return lines

    for full in msgs:
        if "." not in full:
            out += render(full, 0) + [""]
    out += [f"service {api_pb.SERVICE_NAME} {{"] + [f"  rpc {m}({a}) returns ({b});" for m, (a, b) in api_pb.METHODS.items()] + ["}"]
    return "\n".join(out) + "\n"


def parse_proto(text):
    """Minimal proto3 reader: messages (nested), enums, fields.  Returns the same two dicts as schema_as_dict plus services."""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    tok = re.findall(r"map\s*<\s*\w+\s*,\s*\w+\s*>|[A-Za-z_][\w.]*|\d+|[{}=;()<>,\[\]]|\"[^\"]*\"", text)
    msgs, enums, rpcs = {}, {}, {}
    i = 0

    def block(prefix):
        nonlocal i
        while i < len(tok) and tok[i] != "}":
            t = tok[i]
            if t == "message":
                name = prefix + tok[i + 1]
                i += 3
                msgs.setdefault(name, {})
                block(name + ".")
                i += 1
            elif t == "enum":
                name = prefix + tok[i + 1]
                i += 3
                vals = {}
                while tok[i] != "}":
                    if tok[i] == "option" or tok[i] == "reserved":
                        while tok[i] != ";":
                            i += 1
                        i += 1
                        continue
                    vals[int(tok[i + 2])] = tok[i]
                    i += 3
                    while tok[i] != ";":
                        i += 1
                    i += 1
                enums[name] = [vals[k] for k in sorted(vals)]
                i += 1
            elif t in ("option", "reserved", "syntax", "package", "import"):
                while tok[i] != ";":
                    i += 1
                i += 1
            elif t == "oneof":
                i += 3
                block(prefix)
                i += 1
            elif t == "service":
                i += 3
                while tok[i] != "}":
                    if tok[i] == "rpc":
                        rpcs[tok[i + 1]] = (tok[i + 3].split(".")[-1], tok[i + 7].split(".")[-1] if tok[i + 6] == "(" else tok[i + 8].split(".")[-1])
                    i += 1
                i += 1
            elif prefix:   # a field of the current message
                rep = False
                if t in ("repeated", "optional"):
                    rep = t == "repeated"
                    i += 1
                    t = tok[i]
                ts = re.sub(r"\s+", "", t) if t.startswith("map") else t.split(".")[-1]
                name, num = tok[i + 1], int(tok[i + 3])
                msgs[prefix[:-1]][name] = (num, ts, rep)
                while tok[i] != ";":
                    i += 1
                i += 1
            else:
                i += 1

    block("")
    return msgs, enums, rpcs


def diff(proto_text):
    want_m, want_e = schema_as_dict()
    got_m, got_e, got_r = parse_proto(proto_text)
    problems = []
    for m, fields in want_m.items():
        if m not in got_m:
            problems.append(f"message {m}: not in the proto")
            continue
        for name, spec in fields.items():
            g = got_m[m].get(name)
            if g is None:
                by_num = [n for n, s in got_m[m].items() if s[0] == spec[0]]
                problems.append(f"{m}.{name} = {spec[0]}: no such field in the proto" + (f" (number {spec[0]} is {by_num[0]!r} there)" if by_num else ""))
            elif g != spec:
                problems.append(f"{m}.{name}: SCHEMA has (number {spec[0]}, {spec[1]}, repeated={spec[2]}), proto has (number {g[0]}, {g[1]}, repeated={g[2]})")
        for name, g in got_m[m].items():
            if name not in fields:
                problems.append(f"{m}.{name} = {g[0]} ({g[1]}): in the proto, not in SCHEMA (unknown fields are skipped on parse, never sent)")
    for e, vals in want_e.items():
        if e not in got_e:
            problems.append(f"enum {e}: not in the proto")
        elif got_e[e] != vals:
            problems.append(f"enum {e}: SCHEMA {vals} != proto {got_e[e]}")
    for meth, (a, b) in api_pb.METHODS.items():
        if meth in got_r and got_r[meth] != (a, b):
            problems.append(f"rpc {meth}: SCHEMA ({a}) -> ({b}), proto {got_r[meth]}")
        elif meth not in got_r:
            problems.append(f"rpc {meth}: not in the proto's services")
    return problems


if __name__ == "__main__":
    if len(sys.argv) == 2 and sys.argv[1] == "--emit":
        sys.stdout.write(emit())
        sys.exit(0)
    if len(sys.argv) != 2:
        sys.exit(__doc__)
    probs = diff(open(sys.argv[1]).read())
    for p in probs:
        print("MISMATCH", p)
    print(f"{len(probs)} mismatch(es) between api_pb.SCHEMA and {sys.argv[1]}")
    sys.exit(1 if probs else 0)
