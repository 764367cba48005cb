require "./device_indexable"

module Phase
  # Row-major N-D array resident in HBM: the device twin of `NArray(T)`. It includes the same
  # mixin `NArray` includes, so every caller of `MultiIndexable` / `MultiWritable` / `View`
  # keeps working; `DeviceIndexable` then replaces the per-element defaults with kernel launches.
  #
  # ```crystal
  # a = NArray.build(8192, 8192) { |c| (c[0] ^ c[1]).to_f32 }.to_device
  # b = DeviceNArray(Float32).fill([8192, 8192], 0.5_f32)
  # c = (a * b + a)[0..2.., ..-1] # two elementwise kernels and one gather, nothing leaves HBM
  # max, coord = c.argmax
  # host = c.to_host # explicit transfer back
  # ```
  class DeviceNArray(T)
    include DeviceIndexable(T) # brings MultiIndexable::Mutable(T) with it

    getter dev : DeviceBuffer
    getter desc : LibPhGpu::Desc
    @shape : Array(Int32)

    # Uninitialised storage of the given shape.
    def initialize(shape : Enumerable(Int))
      @shape = shape.map do |dim|
        raise DimensionError.new("Cannot create DeviceNArray: One or more of the provided dimensions was negative.") if dim < 0
        dim.to_i32
      end.to_a
      @desc = Descriptor.contiguous(@shape)
      @dev = DeviceBuffer.new(Descriptor.element_count(@shape) * sizeof(T))
    end

    # Aliases an existing buffer (`reshape`).
    protected def initialize(shape : Array(Int32), @dev : DeviceBuffer)
      @shape = shape.dup
      @desc = Descriptor.contiguous(@shape)
    end

    protected def shape_internal : Array(Int32)
      @shape
    end

    # An array over storage the caller provides (the peer-mapped shard of a `ShardedNArray`).
    def self.over(shape : Enumerable(Int), dev : DeviceBuffer) : self
      new(shape.map(&.to_i32).to_a, dev)
    end

    # Picked up by `MultiIndexable#map_with` through `responds_to?`; a host Slice is uploaded.
    def self.of_buffer(shape : Array(Int32), buffer : Slice(T)) : self
      from_host(NArray.of_buffer(shape, buffer))
    end

    # `NArray#to_device`: the explicit host -> device transfer.
    def self.from_host(src : NArray(T)) : self
      result = new(src.shape)
      if result.size > 0
        Device.check LibPhGpu.ph_h2d(result.dev.ptr, src.buffer.to_unsafe.as(Void*), LibC::SizeT.new(src.buffer.bytesize))
        Device.wait # the GC may free `src` as soon as we return
      end
      result
    end

    def self.fill(shape : Enumerable(Int), value : T) : self
      result = new(shape)
      result.fill!(value)
      result
    end

    def self.build(*args, **opts, &block)
      raise DeviceBlockError.new("build")
    end

    def fill!(value : T) : self
      if size > 0
        d = @desc
        Device.check LibPhGpu.ph_fill_region(sizeof(T).to_i32, @dev.ptr, pointerof(d), pointerof(value).as(Void*))
      end
      self
    end

    def zero! : self
      fill!(T.zero)
    end

    # Deep copy, like `NArray#clone`.
    def clone : self
      result = DeviceNArray(T).new(@shape)
      Device.check LibPhGpu.ph_d2d(result.dev.ptr, @dev.ptr, LibC::SizeT.new(size * sizeof(T))) if size > 0
      result
    end

    def dup : self
      clone
    end

    # Like `NArray#reshape`, the result shares this array's buffer.
    def reshape(new_shape : Enumerable(Int)) : self
      new_shape = new_shape.map(&.to_i32).to_a
      if Descriptor.element_count(new_shape) != size
        raise ShapeError.new("Cannot change shape from #{@shape.join('x')} (#{size} elements) to #{new_shape.join('x')} (#{Descriptor.element_count(new_shape)} elements) because reshape cannot add or remove elements.")
      end
      DeviceNArray(T).new(new_shape, @dev)
    end

    def reshape(first : Int, *rest : Int) : self
      reshape([first.to_i32] + rest.map(&.to_i32).to_a)
    end

    def flatten : self
      reshape([size.to_i32])
    end

    # Copying transforms = view + `to_narr`, as in `MultiIndexable#permute / #reverse`.
    def permute(order : Enumerable? = nil) : DeviceNArray(T)
      view.permute(order).to_narr
    end

    def permute(first : Int, *rest : Int) : DeviceNArray(T)
      permute([first.to_i32] + rest.map(&.to_i32).to_a)
    end

    def reverse : DeviceNArray(T)
      view.reverse.to_narr
    end

    # ---- joins (`src/n_array.cr:321-344, 666-750`): one strided copy per input ---------------

    # `NArray#compatible?` (`src/n_array.cr:666-673`), kept as it is: the comparison is `idx != axis`
    # on the raw argument, so a negative axis excludes nothing.
    def compatible?(*others : DeviceIndexable(T), axis = -1) : Bool
      DeviceNArray.compatible_shapes?(shape, others.map(&.shape).to_a, axis)
    end

    def self.compatible_shapes?(first : Array(Int32), others : Array(Array(Int32)), axis : Int) : Bool
      first.each_with_index do |dim, idx|
        others.each do |other|
          return false if dim != other[idx] && idx != axis # `other[idx]` raises IndexError on a shorter shape
        end
      end
      true
    end

    # `NArray.concatenate(*narrs, axis)`: the inputs side by side along `axis`; arrays and views alike.
    def self.concatenate(*narrs : DeviceIndexable(T), axis = 0) : DeviceNArray(T)
      first = narrs[0]
      shapes = narrs.map(&.shape).to_a
      unless compatible_shapes?(first.shape, shapes, axis) && shapes.all? { |other| other.size == first.shape.size }
        raise DimensionError.new("Cannot concatenate these arrays along axis #{axis}: shapes do not match")
      end
      concat_shape = first.shape
      concat_shape[axis] = narrs.sum { |narr| narr.shape[axis] } # IndexError when `axis` is not an axis
      ax = axis < 0 ? axis + concat_shape.size : axis
      result = DeviceNArray(T).new(concat_shape)
      at = 0
      narrs.each do |narr|
        n = narr.shape[ax]
        if n > 0 && result.size > 0
          literal = Array(Range(Int32, Int32)).new(concat_shape.size) { |i| i == ax ? (at..at + n - 1) : (0..concat_shape[i] - 1) }
          result.unsafe_set_chunk(IndexRegion.new(literal, concat_shape, drop: false), narr)
        end
        at += n
      end
      result
    end

    def concatenate(*others : DeviceIndexable(T), axis = 0) : DeviceNArray(T)
      DeviceNArray(T).concatenate(self, *others, axis: axis)
    end

    # `NArray#push`, in place: the buffers are appended as they lie and only `shape[0]` grows,
    # whatever `axis` says (`axis` merely relaxes the compatibility test -- the reference's own
    # TODO). Arrays made by `reshape` before the push keep the old buffer.
    def push(*others : DeviceIndexable(T), axis = 0) : self
      raise DimensionError.new("Cannot concatenate these arrays along axis #{axis}: shapes do not match") if !compatible?(*others, axis: axis)
      total = size + others.sum(&.size)
      grown = DeviceBuffer.new(total * sizeof(T))
      at = append_flat(self, grown, 0_i64)
      others.each { |narr| at = append_flat(narr, grown, at) }
      @shape[0] += others.sum { |narr| narr.shape[0] }
      @dev = grown
      @desc = Descriptor.contiguous(@shape)
      self
    end

    def <<(other : DeviceIndexable(T)) : self
      push(other)
    end

    # Copies `narr`'s elements (lexicographic order) to element `at` of `grown`; returns the next free element.
    private def append_flat(narr : DeviceIndexable(T), grown : DeviceBuffer, at : Int64) : Int64
      flat = narr.is_a?(DeviceNArray(T)) ? narr : narr.to_narr
      if flat.size > 0
        dst = (grown.ptr.as(UInt8*) + at * sizeof(T)).as(Void*)
        Device.check LibPhGpu.ph_d2d(dst, flat.dev.ptr, LibC::SizeT.new(flat.size * sizeof(T)))
      end
      at + flat.size
    end

    # `NArray.wrap(*objects, pad: false)`: a new leading axis with one input per row.
    def self.wrap(*objects : DeviceIndexable(T), pad = false) : DeviceNArray(T)
      raise NotImplementedError.new("As of this time, NArray.wrap() cannot pad arrays for you.") if pad
      container = objects[0].shape
      if objects.any? { |obj| obj.shape != container }
        raise DimensionError.new("Cannot wrap these arrays: shapes do not match. Pass argument pad:true if you want to reshape arrays as necessary.")
      end
      rows = objects.map { |obj| obj.to_narr.reshape([1] + container) }
      concatenate(*rows, axis: 0)
    end

    # Releases the HBM now instead of at the next GC cycle.
    def free : Nil
      @dev.free
    end
  end

  class NArray(T)
    # The explicit host -> device transfer.
    def to_device : DeviceNArray(T)
      DeviceNArray(T).from_host(self)
    end
  end
end
