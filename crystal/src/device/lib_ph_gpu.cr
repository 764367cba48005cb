# lib_ph_gpu.cr -- Crystal `lib` binding of libphgpu.so (include/ph_gpu.h), one `fun` per C entry.
# UNCOMPILED in this repository's build image (no Crystal toolchain there); it is kept in step
# with the header by tests/test_crystal_binding.py, which parses both and compares every
# function name, arity and enum value.
@[Link(ldflags: "-L#{__DIR__}/../../../ph-core_b200/lib -lphgpu -Wl,-rpath,#{__DIR__}/../../../ph-core_b200/lib")]
lib LibPhGpu
  MAX_RANK = 8

  # struct ph_desc: what an IndexRegion or a View transform chain compiles to (element units,
  # signed strides, stride 0 = broadcast axis)
  struct Desc
    rank : Int32
    _pad : Int32
    offset : Int64
    extent : StaticArray(Int64, 8)
    stride : StaticArray(Int64, 8)
  end

  enum Status : Int32
    Ok          = 0
    ErrCuda     = 1
    ErrInvalid  = 2
    ErrUnsupported = 3
    ErrNccl     = 4
    ErrNotInit  = 5
  end

  enum DType : Int32
    F32 = 0
    F64 = 1
    I32 = 2
    I64 = 3
    U8  = 4 # UInt8 and Bool
    I8  = 5
    I16 = 6
    U16 = 7
    U32 = 8
    U64 = 9
  end

  enum Op : Int32
    Add      =  0
    Sub      =  1
    Mul      =  2
    Div      =  3
    FloorDiv =  4
    Mod      =  5
    Pow      =  6
    WAdd     =  7
    WSub     =  8
    WMul     =  9
    WPow     = 10
    And      = 11
    Or       = 12
    Xor      = 13
    Powi     = 14
  end

  enum Cmp : Int32
    Gt = 0
    Lt = 1
    Ge = 2
    Le = 3
    Eq = 4
    Ne = 5
  end

  enum Unary : Int32
    Pos = 0
    Neg = 1
    Not = 2
  end

  enum Red : Int32
    Sum    = 0
    Min    = 1
    Max    = 2
    ArgMax = 3
    ArgMin = 4
  end

  enum HeatMode : Int32
    Fixed     = 0
    Example1D = 1
  end

  FLAG_OVERFLOW = 1_u32
  FLAG_DIV0     = 2_u32
  FLAG_NAN      = 4_u32
  FLAG_ARGUMENT = 8_u32

  # ---- runtime / storage
  fun ph_init(device : Int32) : Int32
  fun ph_shutdown : Int32
  fun ph_device_count(out_count : Int32*) : Int32
  fun ph_sm_count(out_count : Int32*) : Int32
  fun ph_alloc(nbytes : LibC::SizeT, out_dev : Void**) : Int32
  fun ph_free(dev : Void*) : Int32
  fun ph_h2d(dst_dev : Void*, src_host : Void*, nbytes : LibC::SizeT) : Int32
  fun ph_d2h(dst_host : Void*, src_dev : Void*, nbytes : LibC::SizeT) : Int32
  fun ph_d2h_flags(dst_host : Void*, src_dev : Void*, nbytes : LibC::SizeT, out_flags : UInt32*) : Int32
  fun ph_d2h_async(dst_host : Void*, src_dev : Void*, nbytes : LibC::SizeT) : Int32
  fun ph_d2d(dst_dev : Void*, src_dev : Void*, nbytes : LibC::SizeT) : Int32
  fun ph_host_alloc(nbytes : LibC::SizeT, out_host : Void**) : Int32
  fun ph_host_free(host : Void*) : Int32
  fun ph_sync : Int32
  fun ph_stream : Void*
  fun ph_set_stream(cuda_stream : Void*) : Int32
  fun ph_stream_create(out_stream : Void**) : Int32
  fun ph_stream_destroy(stream : Void*) : Int32
  fun ph_stream_wait(waiter_stream : Void*, signaler_stream : Void*) : Int32
  fun ph_stream_sync(stream : Void*) : Int32
  fun ph_free_on(dev : Void*, stream : Void*) : Int32
  fun ph_checksum64(dev : Void*, nbytes : LibC::SizeT, word_offset : UInt64, out_host : UInt64*) : Int32
  fun ph_last_error_string : LibC::Char*
  fun ph_take_arith_flags(out_flags : UInt32*) : Int32
  fun ph_timer_start : Int32
  fun ph_timer_stop(out_ms : Float32*) : Int32
  fun ph_launch_count : Int64

  # ---- elementwise / compare / mask
  fun ph_ewise_binary(op : Int32, dtype : Int32, a : Void*, a_desc : Desc*, b : Void*, b_desc : Desc*,
                      result : Void*, out_desc : Desc*) : Int32
  fun ph_ewise_scalar(op : Int32, dtype : Int32, a : Void*, a_desc : Desc*, scalar_host : Void*,
                      scalar_on_left : Int32, result : Void*, out_desc : Desc*) : Int32
  fun ph_ewise_unary(op : Int32, dtype : Int32, a : Void*, a_desc : Desc*, result : Void*, out_desc : Desc*) : Int32
  fun ph_ewise_mul_add(dtype : Int32, a : Void*, a_desc : Desc*, b : Void*, b_desc : Desc*,
                       c : Void*, c_desc : Desc*, result : Void*, out_desc : Desc*) : Int32
  fun ph_compare(cmp : Int32, dtype : Int32, a : Void*, a_desc : Desc*, b : Void*, b_desc : Desc*,
                 result : UInt8*, out_desc : Desc*) : Int32
  fun ph_compare_scalar(cmp : Int32, dtype : Int32, a : Void*, a_desc : Desc*, scalar_host : Void*,
                        scalar_on_left : Int32, result : UInt8*, out_desc : Desc*) : Int32
  fun ph_compare3(dtype : Int32, a : Void*, a_desc : Desc*, b : Void*, b_desc : Desc*, result : Int32*, out_desc : Desc*) : Int32
  fun ph_compare3_scalar(dtype : Int32, a : Void*, a_desc : Desc*, scalar_host : Void*, scalar_on_left : Int32,
                         result : Int32*, out_desc : Desc*) : Int32
  fun ph_mask_set_scalar(elem_size : Int32, dst : Void*, dst_desc : Desc*, mask : UInt8*, mask_desc : Desc*,
                         scalar_host : Void*) : Int32
  fun ph_mask_set_array(elem_size : Int32, dst : Void*, dst_desc : Desc*, mask : UInt8*, mask_desc : Desc*,
                        src : Void*, src_desc : Desc*) : Int32

  # ---- gather / scatter / fill
  fun ph_copy_strided(elem_size : Int32, src : Void*, src_desc : Desc*, dst : Void*, dst_desc : Desc*) : Int32
  fun ph_fill_region(elem_size : Int32, dst : Void*, dst_desc : Desc*, scalar_host : Void*) : Int32

  # ---- reductions
  fun ph_reduce_full(red : Int32, dtype : Int32, a : Void*, a_desc : Desc*, out_value_host : Void*,
                     out_index_host : Int64*) : Int32
  fun ph_reduce_full_dev(red : Int32, dtype : Int32, a : Void*, a_desc : Desc*, out_value_dev : Void*,
                         out_index_dev : Int64*) : Int32
  fun ph_reduce_axis(red : Int32, dtype : Int32, a : Void*, a_desc : Desc*, axis : Int32, result : Void*,
                     out_desc : Desc*) : Int32

  # ---- heat stencil
  fun ph_heat_step(dtype : Int32, rank : Int32, extents : Int64*, coeff_host : Void*, boundary_mode : Int32,
                   input : Void*, output : Void*) : Int32
  fun ph_heat_run(dtype : Int32, rank : Int32, extents : Int64*, coeff_host : Void*, boundary_mode : Int32,
                  buf_a : Void*, buf_b : Void*, steps : Int64, final_is_b : Int32*) : Int32
  fun ph_heat_step_slab(dtype : Int32, rank : Int32, extents : Int64*, coeff_host : Void*, has_lo : Int32,
                        has_hi : Int32, p_begin : Int64, p_end : Int64, input : Void*, output : Void*,
                        cuda_stream : Void*) : Int32
  fun ph_heat_pass_slab(dtype : Int32, rank : Int32, extents : Int64*, coeff_host : Void*, ghost_planes : Int32,
                        two_steps : Int32, has_lo : Int32, has_hi : Int32, p_begin : Int64, p_end : Int64,
                        input : Void*, output : Void*, cuda_stream : Void*) : Int32

  # ---- multi-GPU (one process per GPU)
  fun ph_comm_unique_id(out128 : UInt8*) : Int32
  fun ph_comm_init(nranks : Int32, rank : Int32, id128 : UInt8*) : Int32
  fun ph_comm_destroy : Int32
  fun ph_comm_p2p_ready(out_ready : Int32*) : Int32
  fun ph_nccl_call_count : Int64
  fun ph_symm_alloc(nbytes : LibC::SizeT, out_dev : Void**) : Int32
  fun ph_symm_free(dev : Void*) : Int32
  fun ph_symm_peer(local_dev : Void*, peer_rank : Int32, out_peer_dev : Void**) : Int32
  fun ph_reduce_full_sharded(red : Int32, dtype : Int32, a : Void*, a_desc : Desc*, elems_before : Int64,
                             out_value_host : Void*, out_index_host : Int64*, out_flags : UInt32*) : Int32
  fun ph_allreduce(red : Int32, dtype : Int32, buf_dev : Void*, count : Int64) : Int32
  fun ph_allgather(send_dev : Void*, recv_dev : Void*, nbytes_per_rank : Int64) : Int32
  fun ph_alltoallv(send_dev : Void**, send_bytes : Int64*, recv_dev : Void**, recv_bytes : Int64*) : Int32
  fun ph_alltoall_strided(elem_size : Int32, src_dev : Void*, src_descs : Desc*, dst_symm : Void*, dst_descs : Desc*) : Int32
  fun ph_halo_exchange(send_lo : Void*, recv_lo : Void*, lo_rank : Int32, send_hi : Void*, recv_hi : Void*,
                       hi_rank : Int32, nbytes : Int64, cuda_stream : Void*) : Int32
  fun ph_heat_run_sharded(dtype : Int32, rank : Int32, local_extents : Int64*, coeff_host : Void*,
                          ghost_planes : Int32, buf_a : Void*, buf_b : Void*, steps : Int64,
                          final_is_b : Int32*) : Int32
end
