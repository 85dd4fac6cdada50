"""SURVEY.md §8a row a10 — the expected steady state of /root/reference/README.md:120-126 (five pods Running, the device
plugin pod READY 2/2) must survive our overlay: deploy/probe-daemonset.patch.yaml is applied (strategic merge: lists of
named objects merge by name, like `kubectl patch --type strategic`) to a minimal DaemonSet of the shape the nvdp chart
renders, and what the reference pins stays true: release nvdp / namespace nvidia (README.md:116), two containers
(README.md:125), runtimeClassName nvidia (values.yaml:4), the kubelet socket directory and the config mount."""
import copy
import os

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strategic_merge(base, patch):
    """dicts merge recursively; lists whose items all carry `name` merge by name (patch order appended); anything else is replaced."""
    if isinstance(base, dict) and isinstance(patch, dict):
        out = copy.deepcopy(base)
        for k, v in patch.items():
            out[k] = strategic_merge(base[k], v) if k in base else copy.deepcopy(v)
        return out
    if isinstance(base, list) and isinstance(patch, list) and all(isinstance(x, dict) and "name" in x for x in base + patch):
        out = [copy.deepcopy(x) for x in base]
        index = {x["name"]: i for i, x in enumerate(out)}
        for item in patch:
            if item["name"] in index:
                out[index[item["name"]]] = strategic_merge(out[index[item["name"]]], item)
            else:
                out.append(copy.deepcopy(item))
        return out
    return copy.deepcopy(patch)


def load(rel):
    with open(os.path.join(ROOT, rel)) as f:
        return yaml.safe_load(f)


def test_probe_patch_keeps_the_steady_state_of_the_reference_install():
    ds = load("tests/golden/daemonset_nvdp_fixture.yaml")
    patched = strategic_merge(ds, load("deploy/probe-daemonset.patch.yaml"))
    # pinned by the reference: release / namespace (README.md:116), pod name prefix and 2/2 (README.md:125)
    assert patched["metadata"]["name"] == "nvdp-nvidia-device-plugin" and patched["metadata"]["namespace"] == "nvidia"
    assert patched["metadata"]["labels"]["app.kubernetes.io/instance"] == "nvdp"
    spec = patched["spec"]["template"]["spec"]
    assert spec["runtimeClassName"] == "nvidia"                                   # values.yaml:4
    names = [c["name"] for c in spec["containers"]]
    assert len(names) == 2 and len(set(names)) == 2, "the plugin pod must stay 2/2 (README.md:125): the patch may not drop or add a container"
    ctr = next(c for c in spec["containers"] if c["name"] == "nvidia-device-plugin-ctr")
    side = next(c for c in spec["containers"] if c["name"] != "nvidia-device-plugin-ctr")
    assert side == next(c for c in ds["spec"]["template"]["spec"]["containers"] if c["name"] == side["name"]), "the second container is untouched"
    # our host replaces the entry point of the plugin container and reads the SAME rendered config (values.yaml:6-18 -> /config/config.yaml)
    assert ctr["command"] == ["/opt/b200probe/bin/b200-device-plugin"] and ctr["args"] == ["--config-file", "/config/config.yaml"]
    mounts = {m["name"]: m["mountPath"] for m in ctr["volumeMounts"]}
    assert mounts["device-plugin"] == "/var/lib/kubelet/device-plugins"            # where kubelet.sock and our socket live
    assert mounts["config"] == "/config"
    assert mounts["features-d"] == "/etc/kubernetes/node-feature-discovery/features.d"      # the label hand-off (values.yaml:1-2 mechanism)
    env = {e["name"]: e["value"] for e in ctr["env"]}
    assert env["CONFIG_FILE"] == "/config/config.yaml" and env["B200PROBE_INTERVAL_S"] == "600"
    vols = {v["name"]: v for v in spec["volumes"]}
    assert set(vols) == {"device-plugin", "available-configs", "config", "features-d"}
    assert vols["features-d"]["hostPath"] == {"path": "/etc/kubernetes/node-feature-discovery/features.d", "type": "DirectoryOrCreate"}
    # nothing of the reference's own files is edited: values.yaml is accepted byte for byte (tests/test_config.py) and the patch adds no chart value
    assert "values" not in load("deploy/probe-daemonset.patch.yaml")


def test_gated_overlays_select_on_the_probe_gate_and_keep_the_reference_requests():
    """configs 3/5: the overlays add ONLY the node selector on the gate label (mirroring the commented selector of
    /root/reference/nvidia-smi.yaml:6-7); runtimeClassName and the nvidia.com/gpu: "1" limit of the reference stay."""
    smi = load("deploy/nvidia-smi-gated.yaml")
    assert smi["spec"]["runtimeClassName"] == "nvidia" and smi["spec"]["restartPolicy"] == "Never"
    assert smi["spec"]["containers"][0]["resources"]["limits"]["nvidia.com/gpu"] == "1"
    assert smi["spec"]["nodeSelector"]["nvidia.com/b200probe.healthy"] == "true"
    jf = load("deploy/jellyfin-gated.patch.yaml")
    assert jf["spec"]["template"]["spec"]["nodeSelector"]["nvidia.com/b200probe.healthy"] == "true"
