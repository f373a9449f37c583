#!/usr/bin/env python
"""Caller-side liveness check, given a cluster — the only contract the reference itself pins for this path:

  * the katib Deployments become available        (/root/reference/testing/kfctl/kf_is_ready_test.py:65-70: `wait_for_deployment`
                                                    over katib-controller / katib-db / katib-manager / katib-ui)
  * a submitted tuning job reaches "Running"       (/root/reference/testing/katib_studyjob_test.py:196-206: `wait_for_condition(...,
                                                    ["Running"])` on the StudyJob; v1beta1 Katib calls the resource Experiment)

and, beyond what the reference checks, that the job's suggestion Deployment (the pod running THIS image) is available and
that at least one trial was created from a suggestion it served.  Uses `kubectl` (no Python Kubernetes client is assumed):

    kubectl apply -f deploy/katib-config-patch.yaml          # register the image for algorithm `bayesianoptimization`
    python deploy/check_live.py --experiment deploy/experiment-example.yaml --namespace kubeflow --timeout 600

Exit code 0 = live.  Never exercised from this repository's CI: there is no cluster here (out of scope)."""
from __future__ import annotations

import argparse
import json
import subprocess
import sys
import time


def kubectl(*args, check=True):
    r = subprocess.run(["kubectl", *args], capture_output=True, text=True)
    if check and r.returncode != 0:
        raise RuntimeError(f"kubectl {' '.join(args)} failed: {r.stderr.strip()}")
    return r.stdout


def deployment_available(namespace, name):
    out = kubectl("-n", namespace, "get", "deployment", name, "-o", "json", check=False)
    if not out:
        return False
    st = json.loads(out).get("status", {})
    return st.get("availableReplicas", 0) >= 1


def experiment_conditions(namespace, name):
    out = kubectl("-n", namespace, "get", "experiments.kubeflow.org", name, "-o", "json", check=False)
    if not out:
        return [], {}
    obj = json.loads(out)
    return [c["type"] for c in obj.get("status", {}).get("conditions", []) if c.get("status") == "True"], obj.get("status", {})


def wait(pred, what, timeout, interval=5.0):
    end = time.time() + timeout
    while time.time() < end:
        if pred():
            print(f"ok: {what}")
            return True
        time.sleep(interval)
    print(f"TIMEOUT: {what}")
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--experiment", required=True, help="Experiment YAML to submit (deploy/experiment-example.yaml)")
    ap.add_argument("--namespace", default="kubeflow")
    ap.add_argument("--timeout", type=float, default=600.0)
    ap.add_argument("--keep", action="store_true", help="leave the Experiment in the cluster")
    a = ap.parse_args()
    ok = True
    for dep in ("katib-controller", "katib-db-manager", "katib-ui"):        # v1beta1 names; the reference's era had katib-manager / katib-db
        ok &= wait(lambda d=dep: deployment_available(a.namespace, d), f"deployment {dep} available", a.timeout / 4)
    name = json.loads(kubectl("apply", "-n", a.namespace, "-f", a.experiment, "-o", "json"))["metadata"]["name"]
    try:
        ok &= wait(lambda: "Running" in experiment_conditions(a.namespace, name)[0], f"experiment {name} reached Running", a.timeout)
        ok &= wait(lambda: deployment_available(a.namespace, f"{name}-bayesianoptimization"), f"suggestion deployment {name}-bayesianoptimization available",
                   a.timeout)
        ok &= wait(lambda: experiment_conditions(a.namespace, name)[1].get("trials", 0) >= 1, "a trial was created from a served suggestion", a.timeout)
    finally:
        if not a.keep:
            kubectl("-n", a.namespace, "delete", "experiments.kubeflow.org", name, "--ignore-not-found", check=False)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
