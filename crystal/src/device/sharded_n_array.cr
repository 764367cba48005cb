require "./device_n_array"
require "./device_view"

module Phase
  # One process per GPU: the communicator of the job and who owns which rows.
  module Comm
    @@world = 1
    @@rank = 0

    def self.world : Int32
      @@world
    end

    def self.rank : Int32
      @@rank
    end

    # Rank 0 creates the id; the launcher hands it to the other ranks (file, environment, MPI ...).
    def self.unique_id : StaticArray(UInt8, 128)
      Device.ensure_init
      id = StaticArray(UInt8, 128).new(0_u8)
      Device.check LibPhGpu.ph_comm_unique_id(id.to_unsafe)
      id
    end

    def self.init(nranks : Int32, rank : Int32, id : StaticArray(UInt8, 128)) : Nil
      Device.ensure_init
      Device.check LibPhGpu.ph_comm_init(nranks, rank, id.to_unsafe)
      @@world, @@rank = nranks, rank
    end

    def self.destroy : Nil
      LibPhGpu.ph_comm_destroy
      @@world, @@rank = 1, 0
    end

    # true: every rank has mapped its peers' memory (CUDA IPC over NVLink): reductions combine inside the
    # kernel and `permute` is one pass of peer stores; false: the NCCL forms of the same entry points run.
    def self.p2p_ready? : Bool
      Device.check LibPhGpu.ph_comm_p2p_ready(out ready)
      ready != 0
    end

    # Contiguous split of `n` leading-axis indices: the first `n % world` ranks get one extra
    # (the rule of ph_shard_range, include/ph_host.h).
    def self.shard_range(n : Int, world : Int32 = @@world, rank : Int32 = @@rank) : {Int32, Int32}
      base, extra = n.to_i32 // world, n.to_i32 % world
      start = rank * base + {rank, extra}.min
      {start, start + base + (rank < extra ? 1 : 0)}
    end

    # Peer-mapped device memory (`ph_symm_alloc`; COLLECTIVE: every rank, same order). Released by
    # `Comm.destroy`, never by a finalizer -- a collective release cannot depend on the GC.
    class SymmBuffer < DeviceBuffer
      def initialize(bytesize : Int64)
        @bytesize = bytesize
        @parent = nil
        @ptr = Pointer(Void).null
        Device.check LibPhGpu.ph_symm_alloc(LibC::SizeT.new({bytesize, 1_i64}.max), pointerof(@ptr))
      end

      def free : Nil
      end
    end
  end

  # An `NArray` distributed along axis 0 over the ranks of the job (SURVEY.md 8(f) f-3). Every rank holds the
  # contiguous row range `Comm.shard_range(shape[0])` as an ordinary `DeviceNArray(T)`:
  #
  # * elementwise operators, comparisons and masked stores are local launches;
  # * `sum` / `min` / `max` / `argmax` are ONE launch per rank with the cross-rank combine inside the kernel
  #   (`ph_reduce_full_sharded`): the same value -- and the same exception -- on every rank;
  # * per-axis folds are local unless the folded axis is the sharded one (`sum(axis: 0)`: the partials are
  #   combined with `ph_allreduce`, checked integer sums with `ph_allgather` + the checked axis-0 fold);
  # * `permute` is one pass of peer stores (`ph_alltoall_strided`: the transpose kernel writes every block
  #   straight into its owner's shard over NVLink), or gathers + `ph_alltoallv` + scatters without P2P.
  #
  # The reference is single-process; this class covers the seam north_star partitions
  # (`src/multi_indexable.cr:30-65` for the array it shards).
  class ShardedNArray(T)
    getter shape : Array(Int32)
    getter local : DeviceNArray(T)
    getter row0 : Int32
    getter row1 : Int32

    def initialize(global_shape : Enumerable(Int), @local : DeviceNArray(T))
      @shape = global_shape.map(&.to_i32).to_a
      raise ShapeError.new("a sharded array needs at least one axis") if @shape.empty?
      @row0, @row1 = Comm.shard_range(@shape[0])
      expected = [@row1 - @row0] + @shape[1..]
      unless @local.shape == expected
        raise ShapeError.new("local shard has shape #{@local.shape}, expected #{expected}")
      end
    end

    # Every rank passes the same host array and keeps its own rows.
    def self.from_global(host : NArray(T)) : self
      row0, row1 = Comm.shard_range(host.shape[0])
      rows = row1 > row0 ? host[row0...row1] : NArray(T).new([0] + host.shape[1..]) { T.zero }
      new(host.shape, DeviceNArray(T).from_host(rows))
    end

    def size : Int64
      Descriptor.element_count(@shape)
    end

    private def row_elems : Int64
      @shape.size > 1 ? Descriptor.element_count(@shape[1..]) : 1_i64
    end

    # The whole array on every rank's host (allgather of the shards, padded to the largest one).
    def to_global : NArray(T)
      mine = @local.to_host
      return mine if Comm.world == 1
      rows_max = (@shape[0] + Comm.world - 1) // Comm.world
      slot = rows_max * row_elems * sizeof(T)
      send = DeviceBuffer.new(slot)
      recv = DeviceBuffer.new(slot * Comm.world)
      Device.check LibPhGpu.ph_d2d(send.ptr, @local.dev.ptr, LibC::SizeT.new(@local.size * sizeof(T))) if @local.size > 0
      Device.check LibPhGpu.ph_allgather(send.ptr, recv.ptr, slot)
      raw = Slice(UInt8).new(slot * Comm.world)
      Device.read_checked(raw.to_unsafe.as(Void*), recv.ptr, LibC::SizeT.new(raw.size))
      buffer = Slice(T).new(size.to_i32) { T.zero }
      Comm.world.times do |r|
        a, b = Comm.shard_range(@shape[0], Comm.world, r)
        count = (b - a) * row_elems
        (raw.to_unsafe + r * slot).as(T*).copy_to(buffer.to_unsafe + a * row_elems, count) if count > 0
      end
      NArray(T).of_buffer(@shape, buffer)
    end

    private def same_shape!(other : ShardedNArray)
      unless other.shape == @shape
        raise ShapeError.new("The shape of this MultiIndexable (#{@shape}) does not match the shape of the one provided (#{other.shape}).")
      end
    end

    private def wrap(local : DeviceNArray(U)) : ShardedNArray(U) forall U
      ShardedNArray(U).new(@shape, local)
    end

    # ---- elementwise / comparisons: local (the operator list of src/multi_indexable.cr:960-981)
    {% for op in %w(+ - * / // % ** &+ &- &* > < >= <=) %}
      def {{op.id}}(other : ShardedNArray(T))
        same_shape!(other)
        wrap(@local {{op.id}} other.local)
      end

      def {{op.id}}(other : T)
        wrap(@local {{op.id}} other)
      end
    {% end %}

    def eq(other : ShardedNArray(T)) : ShardedNArray(Bool)
      same_shape!(other)
      wrap(@local.eq(other.local))
    end

    # `narr[mask] = value` on the distributed array (src/n_array.cr:510-551)
    def []=(mask : ShardedNArray(Bool), value : T)
      @local[mask.local] = value
    end

    def []=(mask : ShardedNArray(Bool), value : ShardedNArray(T))
      same_shape!(value)
      @local[mask.local] = value.local
    end

    # ---- full reductions: collective, one launch per rank
    private def reduce_full(red : LibPhGpu::Red) : {T, Int64}
      cell = StaticArray(UInt64, 2).new(0_u64)
      index = -1_i64
      flags = 0_u32
      d = @local.desc
      Device.check LibPhGpu.ph_reduce_full_sharded(red.value, Device.dtype(T), @local.dev.ptr, pointerof(d), @row0.to_i64 * row_elems,
        cell.to_unsafe.as(Void*), pointerof(index), pointerof(flags))
      Device.raise_for(flags)
      {cell.to_unsafe.as(T*).value, index}
    end

    private def need(result : {T, Int64}) : {T, Int64}
      raise Enumerable::EmptyError.new if result[1] < 0
      result
    end

    def sum : T
      reduce_full(LibPhGpu::Red::Sum)[0]
    end

    def min : T
      need(reduce_full(LibPhGpu::Red::Min))[0]
    end

    def max : T
      need(reduce_full(LibPhGpu::Red::Max))[0]
    end

    # README.md:56-61 across shards: {max, coordinate of the FIRST maximum of the global array}
    def argmax : {T, Array(Int32)}
      value, index = need(reduce_full(LibPhGpu::Red::ArgMax))
      {value, index_to_coord(index)}
    end

    def argmin : {T, Array(Int32)}
      value, index = need(reduce_full(LibPhGpu::Red::ArgMin))
      {value, index_to_coord(index)}
    end

    def index_to_coord(index : Int64) : Array(Int32)
      coord = Array(Int32).new(@shape.size, 0)
      (@shape.size - 1).downto(0) do |i|
        coord[i] = (index % @shape[i]).to_i32
        index //= @shape[i]
      end
      coord
    end

    # ---- per-axis folds. axis >= 1: axis 0 survives, the result is still sharded; axis 0: the partial over my
    # rows is combined across ranks and the REPLICATED `DeviceNArray` of shape[1..] comes back on every rank.
    {% for name, red in {sum: "Sum", min: "Min", max: "Max"} %}
      def {{name.id}}(*, axis : Int32)
        unless 0 <= axis < @shape.size
          raise IndexError.new("axis #{axis} is not present in a #{@shape.size}-dimensional MultiIndexable")
        end
        if axis > 0
          kept = @shape.dup
          kept.delete_at(axis)
          return ShardedNArray(T).new(kept, @local.{{name.id}}(axis: axis))
        end
        over_shards(LibPhGpu::Red::{{red.id}})
      end
    {% end %}

    private def over_shards(red : LibPhGpu::Red) : DeviceNArray(T)
      # every decision that can raise is taken on the GLOBAL shape: all ranks reach the collective, or none
      raise Enumerable::EmptyError.new if @shape[0] == 0 && !red.sum?
      out_shape = @shape.size > 1 ? @shape[1..] : [1]
      part = if @row1 > @row0
               case red
               when .sum? then @local.sum(axis: 0)
               when .max? then @local.max(axis: 0)
               else            @local.min(axis: 0)
               end
             else # an empty shard contributes the identity
               DeviceNArray(T).fill(out_shape, red.sum? ? T.zero : (red.max? ? lowest : highest))
             end
      return part if Comm.world == 1
      {% if T < Int %}
        if red.sum? && !Comm.p2p_ready? # (with P2P `ph_allreduce` folds in rank order with checked adds)
          # checked integer sums: an ncclSum would wrap silently. The per-rank partials ([world, inner], rank
          # order = row order) are gathered and folded by the checked axis-0 sum.
          gathered = DeviceNArray(T).new([Comm.world] + out_shape)
          Device.check LibPhGpu.ph_allgather(part.dev.ptr, gathered.dev.ptr, part.size * sizeof(T))
          return gathered.sum(axis: 0)
        end
      {% end %}
      Device.check LibPhGpu.ph_allreduce(red.value, Device.dtype(T), part.dev.ptr, part.size)
      Device.raise_pending # per-axis folds are raise points: the cross-rank fold too
      part
    end

    private def lowest : T
      {% if T < Float %} -T::INFINITY {% else %} T::MIN {% end %}
    end

    private def highest : T
      {% if T < Float %} T::INFINITY {% else %} T::MAX {% end %}
    end

    # ---- `narr[region_literal]` (gather, src/multi_indexable.cr:338-356) across shards. A literal that leaves axis 0
    # whole is local. Anything else re-splits the result over the ranks along ITS axis 0: every (source, destination)
    # block is ONE strided descriptor over the source's rows -- an arithmetic progression -- and a contiguous row
    # range of the destination, stored straight into its owner over NVLink (`ph_alltoall_strided`). `SlicePlan`
    # below is the Crystal statement of `ph_slice_plan_of` (include/ph_host.h), the plan the C++ and Python layers
    # call; they also carry the form without P2P (blocks gathered, `ph_alltoallv`, received in place).
    def [](*literal) : ShardedNArray(T)
      region = IndexRegion.new(literal.to_a, @shape)
      world, me = Comm.world, Comm.rank
      plan = SlicePlan.new(@shape, region, world, me) # send / land descriptors, recv ranges
      if plan.local?
        rest = [..] + literal.to_a[1..]
        rows = @row1 > @row0 ? @local[rest] : DeviceNArray(T).new([0] + region.shape[1..])
        return ShardedNArray(T).new(region.shape, rows)
      end
      new_shape = plan.new_shape
      return ShardedNArray(T).new(new_shape, @local[literal.to_a]) if world == 1 # one rank owns everything
      m0, m1 = Comm.shard_range(new_shape[0])
      my_shape = [m1 - m0] + new_shape[1..]
      unless Comm.p2p_ready?
        raise RuntimeError.new("slicing the sharded axis needs peer-mapped memory in this layer (the Python and C++ layers carry the ph_alltoallv form)")
      end
      result = DeviceNArray(T).over(my_shape, Comm::SymmBuffer.new(Descriptor.element_count(my_shape) * sizeof(T))) # collective
      sources = plan.send
      targets = plan.land
      Device.check LibPhGpu.ph_alltoall_strided(sizeof(T).to_i32, @local.dev.ptr, sources.to_unsafe, result.dev.ptr, targets.to_unsafe)
      ShardedNArray(T).new(new_shape, result)
    end

    # `narr[region_literal] = value` across shards (scatter / fill, src/multi_writable.cr:55-84). A scalar fills this
    # rank's cells of the region (no exchange). A `ShardedNArray` of the region's shape is the gather run backwards
    # with the same plan: the rows of `value` this rank holds leave as contiguous blocks (`ph_alltoallv`), and every
    # block received is scattered into the arithmetic progression of local rows it belongs to (one strided copy).
    def []=(*literal, value : T)
      args = literal.to_a
      region = IndexRegion.new(args, @shape)
      plan = SlicePlan.new(@shape, region, Comm.world, Comm.rank)
      if plan.local?
        @local[[..] + args[1..]] = value if @row1 > @row0
        return value
      end
      plan.send.each do |cells|
        next if Descriptor.count(cells) == 0
        d = cells
        Device.check LibPhGpu.ph_fill_region(sizeof(T).to_i32, @local.dev.ptr, pointerof(d), pointerof(value).as(Void*))
      end
      Device.wait # `value` is a stack temporary
      value
    end

    def []=(*literal, value : ShardedNArray(T))
      args = literal.to_a
      region = IndexRegion.new(args, @shape)
      world = Comm.world
      plan = SlicePlan.new(@shape, region, world, Comm.rank)
      unless value.shape == plan.new_shape
        raise ShapeError.new("Cannot substitute: the given array has shape #{value.shape}, but the region has shape #{plan.new_shape}.")
      end
      if plan.local?
        @local[[..] + args[1..]] = value.local if @row1 > @row0
        return value
      end
      m0, _ = Comm.shard_range(plan.new_shape[0])
      row = plan.new_shape.size > 1 ? Descriptor.element_count(plan.new_shape[1..]) : 1_i64
      incoming = Array(DeviceNArray(T)?).new(world, nil)
      send_ptr = Array(Void*).new(world, Pointer(Void).null)
      recv_ptr = Array(Void*).new(world, Pointer(Void).null)
      send_bytes = Array(Int64).new(world, 0_i64)
      recv_bytes = Array(Int64).new(world, 0_i64)
      world.times do |q|
        lo, hi = plan.recv[q] # rows of `value` I hold that q's shard receives
        if hi > lo && row > 0
          send_ptr[q] = (value.local.dev.ptr.as(UInt8*) + (lo - m0) * row * sizeof(T)).as(Void*)
          send_bytes[q] = (hi - lo) * row * sizeof(T)
        end
        cells = plan.send[q] # where q's rows land in MY shard
        if (count = Descriptor.count(cells)) > 0
          extents = cells.extent
          block = DeviceNArray(T).new(Array(Int32).new(cells.rank) { |i| extents[i].to_i32 })
          incoming[q] = block
          recv_ptr[q] = block.dev.ptr
          recv_bytes[q] = count * sizeof(T)
        end
      end
      Device.check LibPhGpu.ph_alltoallv(send_ptr.to_unsafe, send_bytes.to_unsafe, recv_ptr.to_unsafe, recv_bytes.to_unsafe)
      world.times do |q|
        if block = incoming[q]
          from, onto = block.desc, plan.send[q]
          Device.check LibPhGpu.ph_copy_strided(sizeof(T).to_i32, block.dev.ptr, pointerof(from), @local.dev.ptr, pointerof(onto))
        end
      end
      Device.wait # the received blocks are released with this scope
      value
    end

    # Host plan of a slice across shards: the Crystal statement of `ph_slice_plan_of` (include/ph_host.h).
    #   send[q] : the block this rank owes q, a strided view of ITS shard (every extent 0: nothing)
    #   land[q] : where it lands in q's shard of the result (a contiguous range of q's rows)
    #   recv[q] : {lo, hi}, the rows of the result (global numbering) q holds for this rank
    struct SlicePlan
      getter new_shape : Array(Int32)
      getter send : Array(LibPhGpu::Desc)
      getter land : Array(LibPhGpu::Desc)
      getter recv : Array({Int64, Int64})
      getter? local : Bool

      @shape : Array(Int32)
      @world : Int32
      @lead : Int32?
      @f0 : Int64
      @s0 : Int64

      def initialize(@shape : Array(Int32), region : IndexRegion, @world : Int32, rank : Int32)
        nd = @shape.size
        @new_shape = region.shape.map(&.to_i32)
        kept = (0...nd).reject { |i| region.degeneracy[i] && region.drop }
        @lead = kept.first? # the result's leading axis; nil: every axis indexed (shape [1])
        @f0, @s0 = region.first[0].to_i64, region.stride[0].to_i64
        rnk = @new_shape.size
        blank = Descriptor.make(rnk, 0_i64, Descriptor.axes, Descriptor.axes)
        @send = Array(LibPhGpu::Desc).new(@world) { blank }
        @land = Array(LibPhGpu::Desc).new(@world) { blank }
        @recv = Array({Int64, Int64}).new(@world) { {0_i64, 0_i64} }
        @local = @lead == 0 && @f0 == 0 && @s0 == 1 && region.proper_shape[0] == @shape[0]
        return if @local
        gstride = Array(Int64).new(nd, 1_i64)
        (nd - 2).downto(0) { |i| gstride[i] = gstride[i + 1] * @shape[i + 1] }
        inner = (1...nd).sum(0_i64) { |i| region.first[i].to_i64 * gstride[i] }
        my0, _ = Comm.shard_range(@shape[0], @world, rank)
        lead = @lead
        @world.times do |q|
          @recv[q] = owned(q, rank)
          lo, hi = owned(rank, q)
          next if hi <= lo
          offset = (lead == 0 ? @f0 + @s0 * lo - my0 : @f0 - my0) * gstride[0] + inner
          offset += region.stride[lead].to_i64 * lo * gstride[lead] if lead && lead > 0
          extent, stride = Descriptor.axes, Descriptor.axes
          if kept.empty?
            extent[0] = 1_i64
            stride[0] = 1_i64
          end
          kept.each_with_index do |axis, d|
            extent[d] = axis == lead ? hi - lo : region.proper_shape[axis].to_i64
            stride[d] = region.stride[axis].to_i64 * gstride[axis]
          end
          @send[q] = Descriptor.make(rnk, offset, extent, stride)
          j0, j1 = Comm.shard_range(@new_shape[0], @world, q)
          whole = Descriptor.contiguous([j1 - j0] + @new_shape[1..])
          lext, lstr = whole.extent, whole.stride
          lext[0] = hi - lo
          @land[q] = Descriptor.make(rnk, (lo - j0) * lstr[0], lext, lstr)
        end
      end

      # rows [lo, hi) of the result's leading axis that `dst` owns and whose data `src` holds
      private def owned(src : Int32, dst : Int32) : {Int64, Int64}
        r0, r1 = Comm.shard_range(@shape[0], @world, src)
        j0, j1 = Comm.shard_range(@new_shape[0], @world, dst)
        none = {0_i64, 0_i64}
        return none if r1 <= r0 || j1 <= j0
        if @lead != 0 # axis 0 is ONE row: its owner has everything
          return r0 <= @f0 < r1 ? {j0.to_i64, j1.to_i64} : none
        end
        if @s0 > 0
          lo = {0_i64, -((-(r0 - @f0)) // @s0)}.max
          hi = r1 - 1 >= @f0 ? (r1 - 1 - @f0) // @s0 + 1 : 0_i64
        else
          t = -@s0
          lo = {0_i64, -((-(@f0 - (r1 - 1))) // t)}.max
          hi = @f0 >= r0 ? (@f0 - r0) // t + 1 : 0_i64
        end
        lo, hi = {lo, j0.to_i64}.max, {hi, j1.to_i64}.min
        hi > lo ? {lo, hi} : none
      end
    end

    # ---- `MultiIndexable#permute` (src/multi_indexable.cr:795-803; no pattern = reversed axes) across shards.
    # The result is sharded along ITS axis 0 (old axis k = pattern[0]). `reuse`: an earlier P2P result of the
    # same shape whose peer-mapped storage receives the new one.
    def permute(pattern : Enumerable(Int)? = nil, reuse : ShardedNArray(T)? = nil) : ShardedNArray(T)
      nd = @shape.size
      pat = pattern ? pattern.map(&.to_i32).to_a : (0...nd).to_a.reverse
      unless pat.size == nd && pat.sort == (0...nd).to_a
        raise IndexError.new("Could not use pattern #{pat} to permute: it is not a permutation of the axes of a #{nd}-dimensional MultiIndexable")
      end
      new_shape = pat.map { |axis| @shape[axis] }
      k = pat[0]
      return ShardedNArray(T).new(new_shape, @local.permute(pat)) if k == 0 # axis 0 stays put: no exchange
      j = pat.index(0).not_nil!                                               # where my rows land
      world, me = Comm.world, Comm.rank
      m0, m1 = Comm.shard_range(new_shape[0])
      my_shape = [m1 - m0] + new_shape[1..]
      # my rows x peer q's slice of old axis k, in q's axis order (a VIEW: nothing is copied yet)
      block = ->(q : Int32) do
        k0, k1 = Comm.shard_range(@shape[k], world, q)
        region = Array(Range(Int32?, Int32?) | Int32).new(nd) { |axis| axis == k ? (k0...k1) : (nil..nil) }
        {k1 - k0, @local.view(region).permute(pat)}
      end
      if world > 1 && Comm.p2p_ready?
        result = if reuse
                   raise ShapeError.new("permute(reuse:): shape #{reuse.shape} is not #{new_shape}") unless reuse.shape == new_shape
                   reuse.local
                 else
                   DeviceNArray(T).over(my_shape, Comm::SymmBuffer.new(Descriptor.element_count(my_shape) * sizeof(T))) # collective
                 end
        sources = Array(LibPhGpu::Desc).new(world) { Descriptor.make(nd, 0_i64, Descriptor.axes, Descriptor.axes) }
        targets = Array(LibPhGpu::Desc).new(world) { Descriptor.make(nd, 0_i64, Descriptor.axes, Descriptor.axes) }
        world.times do |q|
          count, view = block.call(q)
          next if count <= 0 || @row1 <= @row0
          sources[q] = view.desc
          q0, q1 = Comm.shard_range(new_shape[0], world, q)
          whole = Descriptor.contiguous([q1 - q0] + new_shape[1..]) # q's shard of the result
          extent, stride = whole.extent, whole.stride
          extent[j] = (@row1 - @row0).to_i64
          targets[q] = Descriptor.make(nd, @row0.to_i64 * stride[j], extent, stride)
        end
        Device.check LibPhGpu.ph_alltoall_strided(sizeof(T).to_i32, @local.dev.ptr, sources.to_unsafe, result.dev.ptr, targets.to_unsafe)
        return reuse || ShardedNArray(T).new(new_shape, result)
      end
      # NCCL form: permuting gathers, personalised all-to-all, scatters
      result = DeviceNArray(T).new(my_shape)
      landing = ->(q : Int32) do
        p0, p1 = Comm.shard_range(@shape[0], world, q)
        Array(Range(Int32?, Int32?) | Int32).new(nd) { |axis| axis == j ? (p0...p1) : (nil..nil) }
      end
      sends = Array(DeviceNArray(T)?).new(world, nil)
      recvs = Array(DeviceNArray(T)?).new(world, nil)
      world.times do |q|
        count, view = block.call(q)
        p0, p1 = Comm.shard_range(@shape[0], world, q)
        if q == me # my own block never leaves the GPU: one permuting copy into the result
          result[landing.call(q)] = view if count > 0 && @row1 > @row0
          next
        end
        sends[q] = view.to_narr if count > 0 && @row1 > @row0
        incoming = new_shape.dup
        incoming[0] = m1 - m0
        incoming[j] = p1 - p0
        recvs[q] = DeviceNArray(T).new(incoming) if Descriptor.element_count(incoming) > 0
      end
      send_ptr = sends.map { |b| b ? b.dev.ptr : Pointer(Void).null }
      recv_ptr = recvs.map { |b| b ? b.dev.ptr : Pointer(Void).null }
      send_bytes = sends.map { |b| b ? b.size * sizeof(T) : 0_i64 }
      recv_bytes = recvs.map { |b| b ? b.size * sizeof(T) : 0_i64 }
      Device.check LibPhGpu.ph_alltoallv(send_ptr.to_unsafe, send_bytes.to_unsafe, recv_ptr.to_unsafe, recv_bytes.to_unsafe)
      world.times do |q|
        if incoming = recvs[q]
          result[landing.call(q)] = incoming
        end
      end
      ShardedNArray(T).new(new_shape, result)
    end

    def permute(first : Int, *rest : Int) : ShardedNArray(T)
      permute([first.to_i32] + rest.map(&.to_i32).to_a)
    end
  end
end
