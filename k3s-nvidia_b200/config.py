"""The plugin configuration document and time-slicing replica semantics.

Source of truth: /root/reference/values.yaml (must stay byte-identical — the drop-in contract):
  :1-2   gfd.enabled: true
  :4     runtimeClassName: nvidia
  :6-8   config.map.default: |-            <- the plugin config document, a YAML string
  :9     version: v1
  :10-11 flags.migStrategy: none
  :12-18 sharing.timeSlicing.{renameByDefault: false, failRequestsGreaterThanOne: false,
         resources: [{name: nvidia.com/gpu, replicas: 4}]}
/root/reference/README.md:112: "line 18 tells the device plugin to treat that one GPU as if it were
actually four GPUs".  Everything about HOW the upstream plugin interprets the document (ID
annotation ``<uuid>::<replica>``, ``.shared`` rename, request-size check) is [RECALLED] — the plugin
is installed un-pinned (README.md:109,116) and not vendored.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import yaml

RESOURCE_PREFIX = "nvidia.com/"
DEFAULT_RESOURCE = "nvidia.com/gpu"
SHARED_SUFFIX = ".shared"
REPLICA_SEP = "::"


class ConfigError(ValueError):
    pass


@dataclass
class ReplicatedResource:
    name: str
    replicas: int
    rename: Optional[str] = None


@dataclass
class TimeSlicing:
    rename_by_default: bool = False
    fail_requests_greater_than_one: bool = False
    resources: List[ReplicatedResource] = field(default_factory=list)


@dataclass
class PluginConfig:
    version: str = "v1"
    mig_strategy: str = "none"
    device_list_strategy: str = "envvar"
    device_id_strategy: str = "uuid"
    pass_device_specs: bool = False
    time_slicing: TimeSlicing = field(default_factory=TimeSlicing)

    def replicated(self, resource: str = DEFAULT_RESOURCE) -> Optional[ReplicatedResource]:
        for r in self.time_slicing.resources:
            if r.name == resource:
                return r
        return None

    def resource_name(self, resource: str = DEFAULT_RESOURCE) -> str:
        """Advertised extended-resource name: unchanged unless renamed (values.yaml:14 false)."""
        r = self.replicated(resource)
        if r is None:
            return resource
        if r.rename:
            return r.rename
        if self.time_slicing.rename_by_default:
            return resource + SHARED_SUFFIX
        return resource

    def replicas(self, resource: str = DEFAULT_RESOURCE) -> int:
        r = self.replicated(resource)
        return r.replicas if r else 1

    def is_shared(self, resource: str = DEFAULT_RESOURCE) -> bool:
        r = self.replicated(resource)
        return r is not None and r.replicas > 1


@dataclass
class HelmValues:
    gfd_enabled: bool
    runtime_class_name: Optional[str]
    configs: Dict[str, PluginConfig]
    raw_configs: Dict[str, str]

    @property
    def default(self) -> PluginConfig:
        if "default" in self.configs:
            return self.configs["default"]
        if len(self.configs) == 1:
            return next(iter(self.configs.values()))
        return PluginConfig()


def _as_bool(v, what):
    if isinstance(v, bool):
        return v
    raise ConfigError(f"{what} must be a boolean, got {v!r}")


def _resource_name(name) -> str:
    if not isinstance(name, str) or not name:
        raise ConfigError(f"resource name must be a non-empty string, got {name!r}")
    if "/" not in name:
        name = RESOURCE_PREFIX + name
    if not name.startswith(RESOURCE_PREFIX):
        raise ConfigError(f"resource name {name!r} must start with {RESOURCE_PREFIX!r}")
    if len(name) > 63:
        raise ConfigError(f"resource name {name!r} longer than 63 characters")
    return name


def parse_plugin_config(text: str) -> PluginConfig:
    """Parse the document of values.yaml:9-18 (``version: v1`` schema)."""
    try:
        doc = yaml.safe_load(text) or {}
    except yaml.YAMLError as e:  # pragma: no cover - message only
        raise ConfigError(f"plugin config is not valid YAML: {e}") from e
    if not isinstance(doc, dict):
        raise ConfigError("plugin config must be a mapping")
    version = doc.get("version")
    if version != "v1":
        raise ConfigError(f"unknown version: {version!r} (expected 'v1')")
    cfg = PluginConfig()
    flags = doc.get("flags") or {}
    if not isinstance(flags, dict):
        raise ConfigError("flags must be a mapping")
    mig = flags.get("migStrategy", "none")
    if mig not in ("none", "single", "mixed"):
        raise ConfigError(f"invalid migStrategy {mig!r}")
    cfg.mig_strategy = mig
    plugin = flags.get("plugin") or {}
    cfg.device_list_strategy = plugin.get("deviceListStrategy", "envvar")
    cfg.device_id_strategy = plugin.get("deviceIDStrategy", "uuid")
    cfg.pass_device_specs = bool(plugin.get("passDeviceSpecs", False))
    if cfg.device_id_strategy not in ("uuid", "index"):
        raise ConfigError(f"invalid deviceIDStrategy {cfg.device_id_strategy!r}")

    sharing = doc.get("sharing") or {}
    ts = sharing.get("timeSlicing") or {}
    if ts:
        t = TimeSlicing()
        if "renameByDefault" in ts:
            t.rename_by_default = _as_bool(ts["renameByDefault"], "renameByDefault")
        if "failRequestsGreaterThanOne" in ts:
            t.fail_requests_greater_than_one = _as_bool(ts["failRequestsGreaterThanOne"], "failRequestsGreaterThanOne")
        res = ts.get("resources") or []
        if not isinstance(res, list):
            raise ConfigError("sharing.timeSlicing.resources must be a list")
        seen = set()
        for r in res:
            if not isinstance(r, dict):
                raise ConfigError("each replicated resource must be a mapping")
            if "name" not in r:
                raise ConfigError("replicated resource is missing a 'name' field")
            if "replicas" not in r:
                raise ConfigError("replicated resource is missing a 'replicas' field")
            name = _resource_name(r["name"])
            rep = r["replicas"]
            if isinstance(rep, bool) or not isinstance(rep, int):
                raise ConfigError(f"replicas must be an integer, got {rep!r}")
            if rep < 1:
                raise ConfigError(f"number of replicas must be >= 1, got {rep}")
            if name in seen:
                raise ConfigError(f"duplicate replicated resource {name!r}")
            seen.add(name)
            rename = r.get("rename")
            if rename is not None:
                rename = _resource_name(rename)
            t.resources.append(ReplicatedResource(name, rep, rename))
        cfg.time_slicing = t
    return cfg


def parse_helm_values(text: str) -> HelmValues:
    """Parse the Helm values file (values.yaml:1-18) exactly as shipped by the reference."""
    doc = yaml.safe_load(text) or {}
    if not isinstance(doc, dict):
        raise ConfigError("values.yaml must be a mapping")
    gfd = bool((doc.get("gfd") or {}).get("enabled", False))
    rcn = doc.get("runtimeClassName")
    cmap = ((doc.get("config") or {}).get("map")) or {}
    raw, parsed = {}, {}
    for k, v in cmap.items():
        if not isinstance(v, str):
            raise ConfigError(f"config.map.{k} must be a YAML string (block scalar)")
        raw[k] = v
        parsed[k] = parse_plugin_config(v)
    return HelmValues(gfd, rcn, parsed, raw)


# ---- replica annotation [RECALLED upstream AnnotatedID] -----------------------------------------
def annotate(uuid: str, replica: int) -> str:
    return f"{uuid}{REPLICA_SEP}{replica}"


def strip_replica(device_id: str) -> str:
    """``GPU-…::3`` -> ``GPU-…``; an unannotated ID is returned unchanged."""
    return device_id.split(REPLICA_SEP, 1)[0]


def has_replica(device_id: str) -> bool:
    return REPLICA_SEP in device_id


def expand_replicas(uuids: List[str], replicas: int) -> List[str]:
    """1 GPU -> ``replicas`` advertised IDs (README.md:112).  replicas == 1 advertises the bare UUID."""
    if replicas <= 1:
        return list(uuids)
    return [annotate(u, r) for u in uuids for r in range(replicas)]
