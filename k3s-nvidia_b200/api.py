"""kubelet device-plugin API ``v1beta1`` — messages and gRPC stubs built at run time.

This is the OUTER drop-in boundary: what kubelet sees from the plugin that
/root/reference/README.md:116 installs.  The ``.proto`` (k8s.io/kubelet/pkg/apis/deviceplugin/
v1beta1/api.proto) is not vendored in the reference and cannot be fetched here, and there is no
``protoc``; the file descriptor below restates it from recall [RECALLED] — package, message and
field NUMBERS are what matter on the wire.  Re-verify against upstream when a network exists.
"""
from __future__ import annotations

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

VERSION = "v1beta1"
HEALTHY = "Healthy"
UNHEALTHY = "Unhealthy"
DEVICE_PLUGIN_PATH = "/var/lib/kubelet/device-plugins/"
KUBELET_SOCKET = DEVICE_PLUGIN_PATH + "kubelet.sock"

_T = descriptor_pb2.FieldDescriptorProto
_STR, _BOOL, _I32, _I64, _MSG = _T.TYPE_STRING, _T.TYPE_BOOL, _T.TYPE_INT32, _T.TYPE_INT64, _T.TYPE_MESSAGE
_OPT, _REP = _T.LABEL_OPTIONAL, _T.LABEL_REPEATED


def _field(msg, name, number, ftype, label=_OPT, type_name=None, json_name=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = ".v1beta1." + type_name
    if json_name:
        f.json_name = json_name
    return f


def _map_entry(parent, entry_name):
    e = parent.nested_type.add()
    e.name = entry_name
    e.options.map_entry = True
    _field(e, "key", 1, _STR)
    _field(e, "value", 2, _STR)


def _build_file() -> descriptor_pb2.FileDescriptorProto:
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto"
    fd.package = "v1beta1"
    fd.syntax = "proto3"

    def msg(name):
        m = fd.message_type.add()
        m.name = name
        return m

    m = msg("DevicePluginOptions")
    _field(m, "pre_start_required", 1, _BOOL)
    _field(m, "get_preferred_allocation_available", 2, _BOOL)

    m = msg("RegisterRequest")
    _field(m, "version", 1, _STR)
    _field(m, "endpoint", 2, _STR)
    _field(m, "resource_name", 3, _STR)
    _field(m, "options", 4, _MSG, type_name="DevicePluginOptions")

    msg("Empty")

    m = msg("ListAndWatchResponse")
    _field(m, "devices", 1, _MSG, _REP, "Device")

    m = msg("TopologyInfo")
    _field(m, "nodes", 1, _MSG, _REP, "NUMANode")

    m = msg("NUMANode")
    _field(m, "ID", 1, _I64)

    m = msg("Device")
    _field(m, "ID", 1, _STR)
    _field(m, "health", 2, _STR)
    _field(m, "topology", 3, _MSG, type_name="TopologyInfo")

    m = msg("PreStartContainerRequest")
    _field(m, "devices_ids", 1, _STR, _REP)
    msg("PreStartContainerResponse")

    m = msg("PreferredAllocationRequest")
    _field(m, "container_requests", 1, _MSG, _REP, "ContainerPreferredAllocationRequest")
    m = msg("ContainerPreferredAllocationRequest")
    _field(m, "available_deviceIDs", 1, _STR, _REP)
    _field(m, "must_include_deviceIDs", 2, _STR, _REP)
    _field(m, "allocation_size", 3, _I32)
    m = msg("PreferredAllocationResponse")
    _field(m, "container_responses", 1, _MSG, _REP, "ContainerPreferredAllocationResponse")
    m = msg("ContainerPreferredAllocationResponse")
    _field(m, "deviceIDs", 1, _STR, _REP)

    m = msg("AllocateRequest")
    _field(m, "container_requests", 1, _MSG, _REP, "ContainerAllocateRequest")
    m = msg("ContainerAllocateRequest")
    _field(m, "devices_ids", 1, _STR, _REP)

    m = msg("CDIDevice")
    _field(m, "name", 1, _STR)

    m = msg("AllocateResponse")
    _field(m, "container_responses", 1, _MSG, _REP, "ContainerAllocateResponse")

    m = msg("ContainerAllocateResponse")
    _map_entry(m, "EnvsEntry")
    _map_entry(m, "AnnotationsEntry")
    _field(m, "envs", 1, _MSG, _REP, "ContainerAllocateResponse.EnvsEntry")
    _field(m, "mounts", 2, _MSG, _REP, "Mount")
    _field(m, "devices", 3, _MSG, _REP, "DeviceSpec")
    _field(m, "annotations", 4, _MSG, _REP, "ContainerAllocateResponse.AnnotationsEntry")
    _field(m, "cdi_devices", 5, _MSG, _REP, "CDIDevice")

    m = msg("Mount")
    _field(m, "container_path", 1, _STR)
    _field(m, "host_path", 2, _STR)
    _field(m, "read_only", 3, _BOOL)

    m = msg("DeviceSpec")
    _field(m, "container_path", 1, _STR)
    _field(m, "host_path", 2, _STR)
    _field(m, "permissions", 3, _STR)

    def service(name, methods):
        s = fd.service.add()
        s.name = name
        for mname, inp, outp, stream in methods:
            me = s.method.add()
            me.name, me.input_type, me.output_type = mname, ".v1beta1." + inp, ".v1beta1." + outp
            me.server_streaming = stream

    service("Registration", [("Register", "RegisterRequest", "Empty", False)])
    service("DevicePlugin", [
        ("GetDevicePluginOptions", "Empty", "DevicePluginOptions", False),
        ("ListAndWatch", "Empty", "ListAndWatchResponse", True),
        ("GetPreferredAllocation", "PreferredAllocationRequest", "PreferredAllocationResponse", False),
        ("Allocate", "AllocateRequest", "AllocateResponse", False),
        ("PreStartContainer", "PreStartContainerRequest", "PreStartContainerResponse", False),
    ])
    return fd


FILE_DESCRIPTOR_PROTO = _build_file()
_pool = descriptor_pool.DescriptorPool()
_file = _pool.Add(FILE_DESCRIPTOR_PROTO)


def _cls(name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName("v1beta1." + name))


DevicePluginOptions = _cls("DevicePluginOptions")
RegisterRequest = _cls("RegisterRequest")
Empty = _cls("Empty")
ListAndWatchResponse = _cls("ListAndWatchResponse")
TopologyInfo = _cls("TopologyInfo")
NUMANode = _cls("NUMANode")
Device = _cls("Device")
PreStartContainerRequest = _cls("PreStartContainerRequest")
PreStartContainerResponse = _cls("PreStartContainerResponse")
PreferredAllocationRequest = _cls("PreferredAllocationRequest")
ContainerPreferredAllocationRequest = _cls("ContainerPreferredAllocationRequest")
PreferredAllocationResponse = _cls("PreferredAllocationResponse")
ContainerPreferredAllocationResponse = _cls("ContainerPreferredAllocationResponse")
AllocateRequest = _cls("AllocateRequest")
ContainerAllocateRequest = _cls("ContainerAllocateRequest")
AllocateResponse = _cls("AllocateResponse")
ContainerAllocateResponse = _cls("ContainerAllocateResponse")
CDIDevice = _cls("CDIDevice")
Mount = _cls("Mount")
DeviceSpec = _cls("DeviceSpec")

_METHODS = {
    "Registration": {"Register": (RegisterRequest, Empty, False)},
    "DevicePlugin": {
        "GetDevicePluginOptions": (Empty, DevicePluginOptions, False),
        "ListAndWatch": (Empty, ListAndWatchResponse, True),
        "GetPreferredAllocation": (PreferredAllocationRequest, PreferredAllocationResponse, False),
        "Allocate": (AllocateRequest, AllocateResponse, False),
        "PreStartContainer": (PreStartContainerRequest, PreStartContainerResponse, False),
    },
}


def add_servicer(server: grpc.Server, service: str, servicer) -> None:
    """Register ``servicer`` (an object with one method per RPC name) as /v1beta1.<service>/…"""
    handlers = {}
    for name, (req, resp, stream) in _METHODS[service].items():
        fn = getattr(servicer, name)
        mk = grpc.unary_stream_rpc_method_handler if stream else grpc.unary_unary_rpc_method_handler
        handlers[name] = mk(fn, request_deserializer=req.FromString, response_serializer=resp.SerializeToString)
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(f"v1beta1.{service}", handlers),))


class _Stub:
    def __init__(self, channel: grpc.Channel, service: str):
        for name, (req, resp, stream) in _METHODS[service].items():
            mk = channel.unary_stream if stream else channel.unary_unary
            setattr(self, name, mk(f"/v1beta1.{service}/{name}", request_serializer=req.SerializeToString,
                                   response_deserializer=resp.FromString))


def RegistrationStub(channel):  # noqa: N802 (gRPC naming)
    return _Stub(channel, "Registration")


def DevicePluginStub(channel):  # noqa: N802
    return _Stub(channel, "DevicePlugin")
