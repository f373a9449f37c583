"""CPU: tools/check_proto.py — the SCHEMA-vs-api.proto differ — agrees with itself on the rendered SCHEMA and reports a
renumbered field, a retyped field, a missing field and a changed enum in an altered proto (no real katib proto is available
offline; this is what a maintainer runs against upstream's file)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("check_proto", os.path.join(ROOT, "tools", "check_proto.py"))
cp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cp)


def test_rendered_schema_round_trips():
    text = cp.emit()
    assert "rpc GetSuggestions(GetSuggestionsRequest) returns (GetSuggestionsReply);" in text
    assert cp.diff(text) == []
    assert cp.diff("// a comment\n/* block */\n" + text.replace("int32 current_request_number = 4;", "int32 current_request_number = 4; // new trials wanted")) == []


def test_differences_are_reported():
    text = cp.emit()
    bad = (text.replace("int32 current_request_number = 4;", "int32 current_request_number = 3;")
               .replace("string step = 4;", "double step = 4;")
               .replace("  repeated string additional_metric_names = 4;\n", "")
               .replace("EARLYSTOPPED = 6;", "EARLYSTOPPED = 6;\n    PAUSED = 8;")
               .replace("string trial_name = 2;", "string trial_name = 2;\n    string extra_field = 9;"))
    probs = "\n".join(cp.diff(bad))
    assert "GetSuggestionsRequest.current_request_number" in probs and "number 3" in probs
    assert "FeasibleSpace.step" in probs and "double" in probs
    assert "ObjectiveSpec.additional_metric_names" in probs and "no such field" in probs
    assert "TrialStatus.TrialConditionType" in probs and "PAUSED" in probs
    assert "extra_field" in probs
