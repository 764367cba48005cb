require "./device"
require "./descriptor"

module Phase
  # What `DeviceNArray` and `DeviceView` share: a device buffer, ONE descriptor and a shape.
  # Including it overrides every `MultiIndexable` / `MultiWritable` method that would otherwise
  # fall to a per-element host loop: gathers, scatters, masked stores, operators, comparisons,
  # reductions become single kernel launches, and every block-taking method raises.
  module DeviceIndexable(T)
    # Included here (not in the classes) so that `DeviceIndexable(T)` is a subtype of
    # `MultiIndexable(T)`: overloads restricted to it are strictly more specific than the
    # mixin's per-element defaults and always win.
    include MultiIndexable::Mutable(T)

    abstract def dev : DeviceBuffer
    abstract def desc : LibPhGpu::Desc

    # A heap copy of the descriptor for the duration of one (synchronous) C call.
    def desc_ptr : LibPhGpu::Desc*
      box = Pointer(LibPhGpu::Desc).malloc(1)
      box.value = desc
      box
    end

    # The reference counts in Int32; device arrays may be larger (2048^3 does not fit).
    def size : Int64
      shape_internal.empty? ? 0_i64 : shape_internal.reduce(1_i64) { |acc, n| acc * n }
    end

    private def elem_size : Int32
      sizeof(T).to_i32
    end

    # ---- blocks are out of scope on the device path --------------------------------------
    {% for name in %w(each fast_each each_coord each_with_coord each_with map map_with map_with_coord map_with_index map! map_with_coord! map_with_index! apply apply! process) %}
      def {{name.id}}(*args, **opts, &block)
        raise DeviceBlockError.new({{name}})
      end

      def {{name.id}}(*args, **opts)
        raise DeviceBlockError.new({{name}})
      end
    {% end %}

    # `Buffered#buffer` on a device array would be an implicit device -> host copy.
    def buffer
      raise DeviceBlockError.new("buffer")
    end

    # ---- single elements: legal but slow (one tiny transfer each) ---------------------------
    def unsafe_fetch_element(coord : Indexable) : T
      value = uninitialized T
      offset = Descriptor.offset_of(desc, coord)
      Device.read_checked(pointerof(value).as(Void*), (dev.ptr.as(T*) + offset).as(Void*), LibC::SizeT.new(sizeof(T)))
      value
    end

    def unsafe_set_element(coord : Indexable, value : T)
      offset = Descriptor.offset_of(desc, coord)
      Device.check LibPhGpu.ph_h2d((dev.ptr.as(T*) + offset).as(Void*), pointerof(value).as(Void*), LibC::SizeT.new(sizeof(T)))
      Device.wait # `value` lives on this stack frame
    end

    # ---- gather: one launch ---------------------------------------------------------------
    def unsafe_fetch_chunk(region : IndexRegion) : DeviceNArray(T)
      result = DeviceNArray(T).new(region.shape)
      if result.size > 0
        src = region.to_descriptor(desc)
        Device.check LibPhGpu.ph_copy_strided(elem_size, dev.ptr, pointerof(src), result.dev.ptr, result.desc_ptr)
      end
      result
    end

    # `View#to_narr` / `MultiIndexable#to_narr`: the view turned into a copy, one gather.
    def to_narr : DeviceNArray(T)
      result = DeviceNArray(T).new(shape_internal)
      if result.size > 0
        src = desc
        Device.check LibPhGpu.ph_copy_strided(elem_size, dev.ptr, pointerof(src), result.dev.ptr, result.desc_ptr)
      end
      result
    end

    # Explicit device -> host transfer into an ordinary `NArray`.
    def to_host : NArray(T)
      flat = to_narr
      slice = Slice(T).new(flat.size.to_i32) # NArray counts in Int32: raises OverflowError beyond that
      Device.read_checked(slice.to_unsafe.as(Void*), flat.dev.ptr, LibC::SizeT.new(slice.bytesize)) # a raise point
      NArray.of_buffer(shape_internal.clone, slice)
    end

    # ---- scatter / fill: one launch --------------------------------------------------------
    def unsafe_set_chunk(region : IndexRegion, src : DeviceIndexable(T))
      dst = region.to_descriptor(desc)
      return if Descriptor.count(dst) == 0
      dst_extent = dst.extent
      region_shape = Array(Int64).new(dst.rank) { |i| dst_extent[i] }
      # `compatible_shapes?` lets trailing ones differ: view the source with the region's extents
      if folded = Descriptor.reshape(src.desc, region_shape)
        src_desc = folded
        Device.check LibPhGpu.ph_copy_strided(elem_size, src.dev.ptr, pointerof(src_desc), dev.ptr, pointerof(dst))
      else
        copy = src.to_narr # a strided source whose extents cannot be regrouped in place
        src_desc = Descriptor.reshape(copy.desc, region_shape).not_nil!
        Device.check LibPhGpu.ph_copy_strided(elem_size, copy.dev.ptr, pointerof(src_desc), dev.ptr, pointerof(dst))
      end
    end

    # A host source is uploaded first (explicitly visible in the signature: it is an NArray).
    def unsafe_set_chunk(region : IndexRegion, src : NArray(T))
      unsafe_set_chunk(region, DeviceNArray(T).from_host(src))
    end

    def unsafe_set_chunk(region : IndexRegion, src : MultiIndexable(T))
      raise DeviceBlockError.new("unsafe_set_chunk from a #{src.class} (only device arrays and NArray can be a source)")
    end

    def unsafe_set_chunk(region : IndexRegion, value : T)
      dst = region.to_descriptor(desc)
      return if Descriptor.count(dst) == 0
      Device.check LibPhGpu.ph_fill_region(elem_size, dev.ptr, pointerof(dst), pointerof(value).as(Void*))
    end

    # ---- masked store -----------------------------------------------------------------------
    def []=(mask : DeviceIndexable(Bool), value : T)
      if mask.shape_internal != shape_internal
        raise DimensionError.new("Cannot perform masking: mask shape does not match array shape.")
      end
      return if size == 0
      d, m = desc, mask.desc
      Device.check LibPhGpu.ph_mask_set_scalar(elem_size, dev.ptr, pointerof(d), mask.dev.ptr.as(UInt8*), pointerof(m), pointerof(value).as(Void*))
    end

    def []=(mask : DeviceIndexable(Bool), value : DeviceIndexable(T))
      if mask.shape_internal != shape_internal
        raise DimensionError.new("Cannot perform masking: mask shape does not match array shape.")
      end
      if value.shape_internal != shape_internal
        raise DimensionError.new("Cannot perform masking: value shape does not match array shape.")
      end
      return if size == 0
      d, m, v = desc, mask.desc, value.desc
      Device.check LibPhGpu.ph_mask_set_array(elem_size, dev.ptr, pointerof(d), mask.dev.ptr.as(UInt8*), pointerof(m), value.dev.ptr, pointerof(v))
    end

    # ---- views: descriptor only, nothing moves -------------------------------------------------
    def view(region = nil) : DeviceView(T)
      v = DeviceView(T).new(dev, desc, shape_internal.clone)
      region ? v.view(region) : v
    end

    def view(first : Int | Range, *rest) : DeviceView(T)
      view([first] + rest.to_a)
    end

    def mutable_view(region = nil) : DeviceView(T)
      view(region)
    end

    def mutable_view(first : Int | Range, *rest) : DeviceView(T)
      view([first] + rest.to_a)
    end

    # ---- elementwise operators: the same list as MultiIndexable's def_elementwise_binary ----
    private def launch_binary(op : LibPhGpu::Op, other : DeviceIndexable(T), result, name : String)
      if shape_internal != other.shape_internal
        if scalar? || other.scalar?
          raise ShapeError.new("The shape of this MultiIndexable (#{shape_internal}) does not match the shape of the one provided (#{other.shape_internal}), so '#{name}' cannot be applied element-wise. Did you mean to call to_scalar on one of the arguments?")
        end
        raise ShapeError.new("The shape of this MultiIndexable (#{shape_internal}) does not match the shape of the one provided (#{other.shape_internal}), so '#{name}' cannot be applied element-wise.")
      end
      if result.size > 0
        a, b = desc, other.desc
        Device.check LibPhGpu.ph_ewise_binary(op.value, Device.dtype(T), dev.ptr, pointerof(a), other.dev.ptr, pointerof(b), result.dev.ptr, result.desc_ptr)
      end
      result
    end

    private def launch_scalar(op : LibPhGpu::Op, scalar : T, on_left : Bool, result)
      if result.size > 0
        a = desc
        Device.check LibPhGpu.ph_ewise_scalar(op.value, Device.dtype(T), dev.ptr, pointerof(a), pointerof(scalar).as(Void*), on_left ? 1 : 0, result.dev.ptr, result.desc_ptr)
      end
      result
    end

    {% for pair in [{"+", "Add"}, {"-", "Sub"}, {"*", "Mul"}, {"//", "FloorDiv"}, {"%", "Mod"}, {"**", "Pow"},
                    {"&+", "WAdd"}, {"&-", "WSub"}, {"&*", "WMul"}, {"&**", "WPow"}, {"&", "And"}, {"|", "Or"}, {"^", "Xor"}] %}
      def {{pair[0].id}}(other : DeviceIndexable(T)) : DeviceNArray(T)
        launch_binary(LibPhGpu::Op::{{pair[1].id}}, other, DeviceNArray(T).new(shape_internal), {{pair[0]}})
      end

      def {{pair[0].id}}(other : T) : DeviceNArray(T)
        launch_scalar(LibPhGpu::Op::{{pair[1].id}}, other, false, DeviceNArray(T).new(shape_internal))
      end

      # `scalar op narr` (src/patches/number.cr) lands here with the operand order kept
      def scalar_on_left_{{pair[1].downcase.id}}(scalar : T) : DeviceNArray(T)
        launch_scalar(LibPhGpu::Op::{{pair[1].id}}, scalar, true, DeviceNArray(T).new(shape_internal))
      end
    {% end %}

    # `/` is the one operator whose result type differs: Int / Int is Float64 in Crystal.
    def /(other : DeviceIndexable(T))
      launch_binary(LibPhGpu::Op::Div, other, DeviceNArray(typeof(T.zero / T.zero)).new(shape_internal), "/")
    end

    def /(other : T)
      launch_scalar(LibPhGpu::Op::Div, other, false, DeviceNArray(typeof(T.zero / T.zero)).new(shape_internal))
    end

    def scalar_on_left_div(scalar : T)
      launch_scalar(LibPhGpu::Op::Div, scalar, true, DeviceNArray(typeof(T.zero / T.zero)).new(shape_internal))
    end

    # `Float ** Int32` is llvm.powi in Crystal: bit-exact on the device (compiler-rt's loop)
    def **(exponent : Int32) : DeviceNArray(T)
      {% if T == Float32 || T == Float64 %}
        result = DeviceNArray(T).new(shape_internal)
        if result.size > 0
          a = desc
          Device.check LibPhGpu.ph_ewise_scalar(LibPhGpu::Op::Powi.value, Device.dtype(T), dev.ptr, pointerof(a), pointerof(exponent).as(Void*), 0, result.dev.ptr, result.desc_ptr)
        end
        result
      {% else %}
        launch_scalar(LibPhGpu::Op::Pow, T.new(exponent), false, DeviceNArray(T).new(shape_internal))
      {% end %}
    end

    {% for pair in [{"+", "Pos"}, {"-", "Neg"}, {"~", "Not"}] %}
      def {{pair[0].id}} : DeviceNArray(T)
        result = DeviceNArray(T).new(shape_internal)
        if result.size > 0
          a = desc
          Device.check LibPhGpu.ph_ewise_unary(LibPhGpu::Unary::{{pair[1].id}}.value, Device.dtype(T), dev.ptr, pointerof(a), result.dev.ptr, result.desc_ptr)
        end
        result
      end
    {% end %}

    private def launch_compare(cmp : LibPhGpu::Cmp, other : DeviceIndexable(T)) : DeviceNArray(Bool)
      result = DeviceNArray(Bool).new(shape_internal)
      if result.size > 0
        a, b = desc, other.desc
        Device.check LibPhGpu.ph_compare(cmp.value, Device.dtype(T), dev.ptr, pointerof(a), other.dev.ptr, pointerof(b), result.dev.ptr.as(UInt8*), result.desc_ptr)
      end
      result
    end

    private def launch_compare(cmp : LibPhGpu::Cmp, scalar : T, on_left = false) : DeviceNArray(Bool)
      result = DeviceNArray(Bool).new(shape_internal)
      if result.size > 0
        a = desc
        Device.check LibPhGpu.ph_compare_scalar(cmp.value, Device.dtype(T), dev.ptr, pointerof(a), pointerof(scalar).as(Void*), on_left ? 1 : 0, result.dev.ptr.as(UInt8*), result.desc_ptr)
      end
      result
    end

    {% for pair in [{">", "Gt"}, {"<", "Lt"}, {">=", "Ge"}, {"<=", "Le"}] %}
      def {{pair[0].id}}(other : DeviceIndexable(T)) : DeviceNArray(Bool)
        if shape_internal != other.shape_internal
          raise ShapeError.new("The shape of this MultiIndexable (#{shape_internal}) does not match the shape of the one provided (#{other.shape_internal}), so '{{pair[0].id}}' cannot be applied element-wise.")
        end
        launch_compare(LibPhGpu::Cmp::{{pair[1].id}}, other)
      end

      def {{pair[0].id}}(other : T) : DeviceNArray(Bool)
        launch_compare(LibPhGpu::Cmp::{{pair[1].id}}, other)
      end
    {% end %}

    # `<=>` of the operator list (multi_indexable.cr:960-981): -1 / 0 / 1 as Int32. Integer element types
    # only: `Float#<=>` is `Int32?` (nil against NaN), which has no device representation.
    def <=>(other : DeviceIndexable(T)) : DeviceNArray(Int32)
      {% if !(T < Int) %}
        {% raise "<=> on the device path is defined for integer element types (Float#<=> is nilable)" %}
      {% end %}
      if shape_internal != other.shape_internal
        raise ShapeError.new("The shape of this MultiIndexable (#{shape_internal}) does not match the shape of the one provided (#{other.shape_internal}), so '<=>' cannot be applied element-wise.")
      end
      result = DeviceNArray(Int32).new(shape_internal)
      if result.size > 0
        a, b = desc, other.desc
        Device.check LibPhGpu.ph_compare3(Device.dtype(T), dev.ptr, pointerof(a), other.dev.ptr, pointerof(b), result.dev.ptr.as(Int32*), result.desc_ptr)
      end
      result
    end

    def <=>(other : T) : DeviceNArray(Int32)
      {% if !(T < Int) %}
        {% raise "<=> on the device path is defined for integer element types (Float#<=> is nilable)" %}
      {% end %}
      result = DeviceNArray(Int32).new(shape_internal)
      if result.size > 0
        a = desc
        Device.check LibPhGpu.ph_compare3_scalar(Device.dtype(T), dev.ptr, pointerof(a), pointerof(other).as(Void*), 0, result.dev.ptr.as(Int32*), result.desc_ptr)
      end
      result
    end

    def eq(other : DeviceIndexable(T)) : DeviceNArray(Bool)
      if shape_internal != other.shape_internal
        raise DimensionError.new("Cannot compute the element-wise equality between a MultiIndexable with shape #{other.shape_internal} and one with shape #{shape_internal}.")
      end
      launch_compare(LibPhGpu::Cmp::Eq, other)
    end

    def eq(value : T) : DeviceNArray(Bool)
      launch_compare(LibPhGpu::Cmp::Eq, value)
    end

    def =~(value : T) : DeviceNArray(Bool)
      eq(value)
    end

    # Printing is a host activity: the reference's `Formatter` walks elements one by one
    # (`multi_indexable/formatter/formatter.cr:61-62` calls `each`), which a device array refuses.
    # The elements come back in ONE explicit transfer and the host array prints itself.
    def to_s(io : IO) : Nil
      io << "device "
      to_host.to_s(io)
    end

    def inspect(io : IO) : Nil
      io << "#<" << self.class.name << " shape=" << shape_internal << ">"
    end

    def to_literal_s(io : IO) : Nil
      to_host.to_literal_s(io)
    end

    def ==(other : DeviceIndexable(T)) : Bool
      return false if shape_internal != other.shape_internal
      return true if size == 0
      eq(other).min
    end

    # NEW: fused `(self * b) + c`, two roundings (never an FMA); b and c may broadcast.
    def mul_add(b : DeviceIndexable(T), c : DeviceIndexable(T)) : DeviceNArray(T)
      result = DeviceNArray(T).new(shape_internal)
      if result.size > 0
        da, db, dc = desc, Descriptor.broadcast(b.desc, shape_internal), Descriptor.broadcast(c.desc, shape_internal)
        Device.check LibPhGpu.ph_ewise_mul_add(Device.dtype(T), dev.ptr, pointerof(da), b.dev.ptr, pointerof(db), c.dev.ptr, pointerof(dc), result.dev.ptr, result.desc_ptr)
      end
      result
    end

    # NEW: `self op other` with size-1 axes stretched (= `tile` + operator in reference terms).
    def broadcast(op : LibPhGpu::Op, other : DeviceIndexable(T)) : DeviceNArray(T)
      shape = ShapeUtil.broadcast_shapes(shape_internal, other.shape_internal)
      result = DeviceNArray(T).new(shape)
      if result.size > 0
        da, db = Descriptor.broadcast(desc, shape), Descriptor.broadcast(other.desc, shape)
        Device.check LibPhGpu.ph_ewise_binary(op.value, Device.dtype(T), dev.ptr, pointerof(da), other.dev.ptr, pointerof(db), result.dev.ptr, result.desc_ptr)
      end
      result
    end

    # ---- reductions: Enumerable's folds, one launch each ------------------------------------------
    private def reduce_full(red : LibPhGpu::Red) : {T, Int64}
      raise Enumerable::EmptyError.new if size == 0
      # record mode (the sharded entry on one process exchanges nothing): one launch, the finishing block writes
      # value, index and the pending flags into a pinned host record -- no copy, no second read for the flags
      cell = StaticArray(UInt64, 2).new(0_u64)
      index = -1_i64
      flags = 0_u32
      a = desc
      Device.check LibPhGpu.ph_reduce_full_sharded(red.value, Device.dtype(T), dev.ptr, pointerof(a), 0_i64,
        cell.to_unsafe.as(Void*), pointerof(index), pointerof(flags))
      Device.raise_for(flags)
      {cell.to_unsafe.as(T*).value, index}
    end

    def sum : T
      return T.zero if size == 0
      reduce_full(LibPhGpu::Red::Sum)[0]
    end

    def min : T
      reduce_full(LibPhGpu::Red::Min)[0]
    end

    def max : T
      reduce_full(LibPhGpu::Red::Max)[0]
    end

    # `{max, coord}` of the FIRST maximum in lexicographic order (the README's each_with_coord idiom).
    def argmax : {T, Array(Int32)}
      value, index = reduce_full(LibPhGpu::Red::ArgMax)
      {value, lex_index_to_coord(index)}
    end

    def argmin : {T, Array(Int32)}
      value, index = reduce_full(LibPhGpu::Red::ArgMin)
      {value, lex_index_to_coord(index)}
    end

    private def lex_index_to_coord(index : Int64) : Array(Int32)
      coord = Array(Int32).new(shape_internal.size, 0)
      (shape_internal.size - 1).downto(0) do |i|
        coord[i] = (index % shape_internal[i]).to_i32
        index //= shape_internal[i]
      end
      coord
    end

    private def reduce_axis(red : LibPhGpu::Red, axis : Int32, result)
      unless 0 <= axis < shape_internal.size
        raise IndexError.new("Axis #{axis} is not present in a #{shape_internal.size}-dimensional MultiIndexable.")
      end
      raise Enumerable::EmptyError.new if shape_internal[axis] == 0 && red != LibPhGpu::Red::Sum
      if result.size > 0
        if shape_internal[axis] == 0
          result.zero! # the sum over no elements
        else
          a = desc
          Device.check LibPhGpu.ph_reduce_axis(red.value, Device.dtype(T), dev.ptr, pointerof(a), axis, result.dev.ptr, result.desc_ptr)
        end
      end
      Device.raise_pending
      result
    end

    private def shape_without(axis : Int32) : Array(Int32)
      rest = [] of Int32
      shape_internal.each_with_index { |n, i| rest << n unless i == axis }
      rest.empty? ? [1] : rest
    end

    # Per-axis forms: `each_slice(axis)` folded with the element-wise operator, index ascending.
    def sum(*, axis : Int32) : DeviceNArray(T)
      reduce_axis(LibPhGpu::Red::Sum, axis, DeviceNArray(T).new(shape_without(axis)))
    end

    def min(*, axis : Int32) : DeviceNArray(T)
      reduce_axis(LibPhGpu::Red::Min, axis, DeviceNArray(T).new(shape_without(axis)))
    end

    def max(*, axis : Int32) : DeviceNArray(T)
      reduce_axis(LibPhGpu::Red::Max, axis, DeviceNArray(T).new(shape_without(axis)))
    end

    def argmax(*, axis : Int32) : DeviceNArray(Int64)
      reduce_axis(LibPhGpu::Red::ArgMax, axis, DeviceNArray(Int64).new(shape_without(axis)))
    end

    def argmin(*, axis : Int32) : DeviceNArray(Int64)
      reduce_axis(LibPhGpu::Red::ArgMin, axis, DeviceNArray(Int64).new(shape_without(axis)))
    end

    # ---- slices / tile ---------------------------------------------------------------------------
    # The reference gathers one chunk per index (ChunkIterator -> unsafe_fetch_chunk); on the device that
    # is one launch per slice and launch-bound for every axis but the leading one. All slices along
    # `axis` together ARE the array with `axis` moved to the front, so ONE permuting copy produces them
    # and they are handed out as consecutive ranges of its buffer (disjoint: still independent arrays).
    def slices(axis = 0) : Array(DeviceNArray(T))
      unless 0 <= axis < shape_internal.size
        raise IndexError.new("Axis #{axis} is not present in a #{shape_internal.size}-dimensional MultiIndexable.")
      end
      count = shape_internal[axis]
      rest = shape_without(axis)
      if count == 0 || size == 0
        return Array(DeviceNArray(T)).new(count) { DeviceNArray(T).new(rest) }
      end
      order = [axis] + (0...shape_internal.size).reject { |i| i == axis }
      moved = view.permute(order).to_narr # one launch
      step = size // count * sizeof(T)
      Array(DeviceNArray(T)).new(count) do |i|
        DeviceNArray(T).new(rest, DeviceBuffer.new(moved.dev, i.to_i64 * step, step))
      end
    end

    # `each_slice` hands out VIEWS over this array's buffer -- descriptors only, no copy and no launch -- so
    # the reference's per-axis idiom (`each_slice(axis) { |s| acc = acc + s }`, multi_indexable.cr:742-786)
    # costs just the consumer's kernels, which read the strided slices directly. `slices` makes independent
    # arrays (one batched copy) for callers that need to own them.
    def each_slice(axis = 0) : Iterator(DeviceView(T))
      unless 0 <= axis < shape_internal.size
        raise IndexError.new("Axis #{axis} is not present in a #{shape_internal.size}-dimensional MultiIndexable.")
      end
      source = desc
      (0...shape_internal[axis]).each.map do |i|
        DeviceView(T).new(dev, Descriptor.drop_axis(source, axis, i.to_i64), shape_without(axis))
      end
    end

    def each_slice(axis = 0, &block : DeviceView(T) ->)
      each_slice(axis).each { |slice| yield slice }
    end

    # `out[c] = self[c % shape]`; as a descriptor every axis becomes (count, extent) with strides
    # (0, stride), so the device never computes a modulo.
    def tile(counts : Enumerable(Int)) : DeviceNArray(T)
      counts = counts.to_a
      raise DimensionError.new("Cannot tile: #{counts.size} counts for #{shape_internal.size} dimensions.") if counts.size != shape_internal.size
      raise ShapeError.new("Cannot tile on the device path beyond #{LibPhGpu::MAX_RANK // 2} dimensions.") if 2 * counts.size > LibPhGpu::MAX_RANK
      src, dst = Descriptor.tile(desc, counts)
      result = DeviceNArray(T).new(shape_internal.map_with_index { |n, i| n * counts[i].to_i32 })
      if result.size > 0
        Device.check LibPhGpu.ph_copy_strided(elem_size, dev.ptr, pointerof(src), result.dev.ptr, pointerof(dst))
      end
      result
    end
  end
end
