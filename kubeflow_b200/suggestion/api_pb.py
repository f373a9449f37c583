"""``api.v1.beta1`` (Katib manager API) messages built programmatically — protoc / grpc_tools are not in this image.

UNVERIFIED FIELD NUMBERS.  kubeflow/katib's ``pkg/apis/manager/v1beta1/api.proto`` is not under /root/reference
 and cannot be fetched; the numbers below are reproduced from memory of upstream and
are kept in ONE table (``SCHEMA``) so a maintainer can diff them against the real proto in a minute.  A wrong number
does not break this repo's own client/server round trips (both sides use this table) but would silently drop fields
against a real katib-controller — INTEGRATION.md says how to regenerate this module from the real proto instead.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

PACKAGE = "api.v1.beta1"
T = descriptor_pb2.FieldDescriptorProto

# message -> [(field name, number, type, label, type_name-or-None)]
_S, _I32, _DBL, _MSG, _ENUM = T.TYPE_STRING, T.TYPE_INT32, T.TYPE_DOUBLE, T.TYPE_MESSAGE, T.TYPE_ENUM
_OPT, _REP = T.LABEL_OPTIONAL, T.LABEL_REPEATED

ENUMS = {
    "ParameterType": ["UNKNOWN_TYPE", "DOUBLE", "INT", "DISCRETE", "CATEGORICAL"],
    "ObjectiveType": ["UNKNOWN", "MINIMIZE", "MAXIMIZE"],
    "ComparisonType": ["UNKNOWN_COMPARISON", "EQUAL", "LESS", "GREATER"],
}
# nested enum: TrialStatus.TrialConditionType
TRIAL_CONDITIONS = ["CREATED", "RUNNING", "SUCCEEDED", "KILLED", "FAILED", "METRICSUNAVAILABLE", "EARLYSTOPPED", "UNKNOWN"]

SCHEMA = {
    "Experiment": [("name", 1, _S, _OPT, None), ("spec", 2, _MSG, _OPT, "ExperimentSpec")],
    "ExperimentSpec": [("parameter_specs", 1, _MSG, _OPT, "ExperimentSpec.ParameterSpecs"), ("objective", 2, _MSG, _OPT, "ObjectiveSpec"),
                       ("algorithm", 3, _MSG, _OPT, "AlgorithmSpec"), ("early_stopping", 4, _MSG, _OPT, "EarlyStoppingSpec"),
                       ("parallel_trial_count", 5, _I32, _OPT, None), ("max_trial_count", 6, _I32, _OPT, None),
                       ("nas_config", 7, _MSG, _OPT, "NasConfig")],
    "ExperimentSpec.ParameterSpecs": [("parameters", 1, _MSG, _REP, "ParameterSpec")],
    "ParameterSpec": [("name", 1, _S, _OPT, None), ("parameter_type", 2, _ENUM, _OPT, "ParameterType"),
                      ("feasible_space", 3, _MSG, _OPT, "FeasibleSpace")],
    "FeasibleSpace": [("max", 1, _S, _OPT, None), ("min", 2, _S, _OPT, None), ("list", 3, _S, _REP, None), ("step", 4, _S, _OPT, None)],
    "ObjectiveSpec": [("type", 1, _ENUM, _OPT, "ObjectiveType"), ("goal", 2, _DBL, _OPT, None), ("objective_metric_name", 3, _S, _OPT, None),
                      ("additional_metric_names", 4, _S, _REP, None)],
    "AlgorithmSpec": [("algorithm_name", 1, _S, _OPT, None), ("algorithm_settings", 2, _MSG, _REP, "AlgorithmSetting")],
    "AlgorithmSetting": [("name", 1, _S, _OPT, None), ("value", 2, _S, _OPT, None)],
    "EarlyStoppingSpec": [("algorithm_name", 1, _S, _OPT, None), ("algorithm_settings", 2, _MSG, _REP, "EarlyStoppingSetting")],
    "EarlyStoppingSetting": [("name", 1, _S, _OPT, None), ("value", 2, _S, _OPT, None)],
    "EarlyStoppingRule": [("name", 1, _S, _OPT, None), ("value", 2, _S, _OPT, None), ("comparison", 3, _ENUM, _OPT, "ComparisonType"),
                          ("start_step", 4, _I32, _OPT, None)],
    "NasConfig": [("graph_config", 1, _MSG, _OPT, "GraphConfig"), ("operations", 2, _MSG, _OPT, "NasConfig.Operations")],
    "NasConfig.Operations": [("operation", 1, _MSG, _REP, "Operation")],
    "GraphConfig": [("num_layers", 1, _I32, _OPT, None), ("input_sizes", 2, _I32, _REP, None), ("output_sizes", 3, _I32, _REP, None)],
    "Operation": [("operation_type", 1, _S, _OPT, None), ("parameter_specs", 2, _MSG, _OPT, "Operation.ParameterSpecs")],
    "Operation.ParameterSpecs": [("parameters", 1, _MSG, _REP, "ParameterSpec")],
    "Trial": [("name", 1, _S, _OPT, None), ("spec", 2, _MSG, _OPT, "TrialSpec"), ("status", 3, _MSG, _OPT, "TrialStatus")],
    "TrialSpec": [("objective", 2, _MSG, _OPT, "ObjectiveSpec"), ("parameter_assignments", 3, _MSG, _OPT, "TrialSpec.ParameterAssignments"),
                  ("labels", 4, "map", _REP, None)],
    "TrialSpec.ParameterAssignments": [("assignments", 1, _MSG, _REP, "ParameterAssignment")],
    "ParameterAssignment": [("name", 1, _S, _OPT, None), ("value", 2, _S, _OPT, None)],
    "TrialStatus": [("start_time", 1, _S, _OPT, None), ("completion_time", 2, _S, _OPT, None),
                    ("condition", 3, _ENUM, _OPT, "TrialStatus.TrialConditionType"), ("observation", 4, _MSG, _OPT, "Observation")],
    "Observation": [("metrics", 1, _MSG, _REP, "Metric")],
    "Metric": [("name", 1, _S, _OPT, None), ("value", 2, _S, _OPT, None)],
    "GetSuggestionsRequest": [("experiment", 1, _MSG, _OPT, "Experiment"), ("trials", 2, _MSG, _REP, "Trial"),
                              ("current_request_number", 4, _I32, _OPT, None), ("total_request_number", 5, _I32, _OPT, None)],
    "GetSuggestionsReply": [("parameter_assignments", 1, _MSG, _REP, "GetSuggestionsReply.ParameterAssignments"),
                            ("algorithm", 2, _MSG, _OPT, "AlgorithmSpec"), ("early_stopping_rules", 3, _MSG, _REP, "EarlyStoppingRule")],
    "GetSuggestionsReply.ParameterAssignments": [("assignments", 1, _MSG, _REP, "ParameterAssignment"), ("trial_name", 2, _S, _OPT, None),
                                                 ("labels", 3, "map", _REP, None)],
    "ValidateAlgorithmSettingsRequest": [("experiment", 1, _MSG, _OPT, "Experiment")],
    "ValidateAlgorithmSettingsReply": [],
}

SERVICE_NAME = "Suggestion"
METHODS = {"GetSuggestions": ("GetSuggestionsRequest", "GetSuggestionsReply"),
           "ValidateAlgorithmSettings": ("ValidateAlgorithmSettingsRequest", "ValidateAlgorithmSettingsReply")}


def _add_fields(msg: descriptor_pb2.DescriptorProto, full: str, fields):
    for name, num, typ, label, tname in fields:
        f = msg.field.add()
        f.name, f.number, f.label = name, num, label
        if typ == "map":  # map<string,string>
            entry = msg.nested_type.add()
            entry.name = "".join(p.capitalize() for p in name.split("_")) + "Entry"
            entry.options.map_entry = True
            for n, k in (("key", 1), ("value", 2)):
                e = entry.field.add()
                e.name, e.number, e.label, e.type = n, k, _OPT, _S
            f.type, f.type_name = _MSG, f".{PACKAGE}.{full}.{entry.name}"
        else:
            f.type = typ
            if tname:
                f.type_name = f".{PACKAGE}.{tname}"


def _build_file() -> descriptor_pb2.FileDescriptorProto:
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "kubeflow_b200/api_v1_beta1.proto", PACKAGE, "proto3"
    for ename, values in ENUMS.items():
        e = fd.enum_type.add()
        e.name = ename
        for i, v in enumerate(values):
            ev = e.value.add()
            ev.name, ev.number = v, i
    top = {}
    for full in SCHEMA:
        if "." not in full:
            m = fd.message_type.add()
            m.name = full
            top[full] = m
    for full, fields in SCHEMA.items():
        if "." in full:
            parent, child = full.split(".")
            m = top[parent].nested_type.add()
            m.name = child
            _add_fields(m, full, fields)
    for full, fields in SCHEMA.items():
        if "." not in full:
            _add_fields(top[full], full, fields)
    e = top["TrialStatus"].enum_type.add()
    e.name = "TrialConditionType"
    for i, v in enumerate(TRIAL_CONDITIONS):
        ev = e.value.add()
        ev.name, ev.number = v, i
    svc = fd.service.add()
    svc.name = SERVICE_NAME
    for mname, (req, rep) in METHODS.items():
        m = svc.method.add()
        m.name, m.input_type, m.output_type = mname, f".{PACKAGE}.{req}", f".{PACKAGE}.{rep}"
    return fd


_POOL = descriptor_pool.DescriptorPool()
FILE_DESCRIPTOR = _POOL.Add(_build_file())


def _cls(name: str):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(f"{PACKAGE}.{name}"))


Experiment = _cls("Experiment")
ExperimentSpec = _cls("ExperimentSpec")
ParameterSpec = _cls("ParameterSpec")
FeasibleSpace = _cls("FeasibleSpace")
ObjectiveSpec = _cls("ObjectiveSpec")
AlgorithmSpec = _cls("AlgorithmSpec")
AlgorithmSetting = _cls("AlgorithmSetting")
EarlyStoppingRule = _cls("EarlyStoppingRule")
Trial = _cls("Trial")
TrialSpec = _cls("TrialSpec")
TrialStatus = _cls("TrialStatus")
Observation = _cls("Observation")
Metric = _cls("Metric")
ParameterAssignment = _cls("ParameterAssignment")
GetSuggestionsRequest = _cls("GetSuggestionsRequest")
GetSuggestionsReply = _cls("GetSuggestionsReply")
ValidateAlgorithmSettingsRequest = _cls("ValidateAlgorithmSettingsRequest")
ValidateAlgorithmSettingsReply = _cls("ValidateAlgorithmSettingsReply")

# enum values (module-level, like the *_pb2 modules protoc would generate)
UNKNOWN_TYPE, DOUBLE, INT, DISCRETE, CATEGORICAL = range(5)
UNKNOWN, MINIMIZE, MAXIMIZE = range(3)
SUCCEEDED = TRIAL_CONDITIONS.index("SUCCEEDED")
EARLYSTOPPED = TRIAL_CONDITIONS.index("EARLYSTOPPED")
FAILED = TRIAL_CONDITIONS.index("FAILED")
KILLED = TRIAL_CONDITIONS.index("KILLED")
METRICSUNAVAILABLE = TRIAL_CONDITIONS.index("METRICSUNAVAILABLE")
