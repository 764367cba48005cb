# Device-resident backing for ph-core's data-parallel hot path (NVIDIA B200, libphgpu.so).
# Add `require "./device"` at the end of src/ph-core.cr (after n_array, view, patches).
require "./device/lib_ph_gpu"
require "./device/device"
require "./device/descriptor"
require "./device/device_indexable"
require "./device/device_n_array"
require "./device/device_view"
require "./device/number_patch"
require "./device/heat"
require "./device/sharded_n_array"
require "./device/pipeline"
