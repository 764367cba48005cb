require "./device_indexable"

module Phase
  # Lazy view of a device array: source buffer + ONE descriptor. `View` composes
  # Region / Permute / Reverse / Reshape coordinate transforms and applies the chain per element
  # (two heap allocations each); on the device the first three are affine in the coordinate, so
  # the chain folds into (offset, extent[], stride[]) as it is built and a read is one gather, a
  # write one scatter. Reshape folds when strides can express it and materialises otherwise.
  # It is both `View` and `MutableView`: writes go through to the source.
  #
  # Every transform is a pure function from one (buffer, descriptor, shape) triple to the next;
  # the `!` forms of the reference's API adopt the transformed triple in place.
  class DeviceView(T)
    include DeviceIndexable(T) # brings MultiIndexable::Mutable(T) with it

    getter dev : DeviceBuffer
    getter desc : LibPhGpu::Desc
    @shape : Array(Int32)

    def initialize(@dev : DeviceBuffer, @desc : LibPhGpu::Desc, @shape : Array(Int32))
    end

    protected def shape_internal : Array(Int32)
      @shape
    end

    private def derive(desc : LibPhGpu::Desc, shape : Array(Int32), dev : DeviceBuffer = @dev) : DeviceView(T)
      DeviceView(T).new(dev, desc, shape)
    end

    protected def adopt(other : DeviceView(T)) : self
      @dev, @desc, @shape = other.dev, other.desc, other.shape_internal
      self
    end

    def clone : self
      derive(@desc, @shape.dup)
    end

    # ---- restriction to a region (`View#view`, `RegionTransform`)
    def view(region = nil) : DeviceView(T)
      case region
      when Nil         then clone
      when IndexRegion then derive(region.to_descriptor(@desc), region.shape)
      else
        canonical = IndexRegion.new(region, @shape)
        derive(canonical.to_descriptor(@desc), canonical.shape)
      end
    end

    def mutable_view(region = nil) : DeviceView(T)
      view(region)
    end

    # A chunk of a view is again a view (`View#unsafe_fetch_chunk`), nothing is copied.
    def unsafe_fetch_chunk(region : IndexRegion) : DeviceView(T)
      view(region)
    end

    # ---- `PermuteTransform`: output axis i is source axis order[i]; no order = reversed axes
    def permute(order : Enumerable? = nil) : DeviceView(T)
      axes = order.try &.map(&.to_i32).to_a
      if axes && (stray = axes.find { |axis| !(0 <= axis < @shape.size) })
        raise IndexError.new("Could not use pattern #{axes} to permute: Axis #{stray} is not present in a #{dimensions}-dimensional MultiIndexable")
      end
      moved = Descriptor.permute(@desc, axes)
      extents = moved.extent
      derive(moved, Array(Int32).new(moved.rank) { |i| extents[i].to_i32 })
    end

    def permute!(order : Enumerable? = nil) : self
      adopt permute(order)
    end

    # ---- `ReverseTransform`: every axis flipped
    def reverse : DeviceView(T)
      derive(Descriptor.reverse(@desc), @shape.dup)
    end

    def reverse! : self
      adopt reverse
    end

    # ---- `ReshapeTransform`: in strides when the new axes subdivide contiguous runs, else through a copy
    def reshape(new_shape : Enumerable) : DeviceView(T)
      target = new_shape.map(&.to_i32).to_a
      wanted = Descriptor.element_count(target)
      if wanted != size
        raise ShapeError.new("Cannot change shape from #{@shape.join('x')} (#{size} elements) to #{target.join('x')} (#{wanted} elements) because reshape cannot add or remove elements.")
      end
      if folded = Descriptor.reshape(@desc, target)
        derive(folded, target)
      else
        copy = to_narr
        derive(Descriptor.contiguous(target), target, copy.dev)
      end
    end

    def reshape!(new_shape : Enumerable) : self
      adopt reshape(new_shape)
    end

    # Splat forms (`view.permute(1, 0, 2)`); at least one argument, so that a bare `permute`
    # always means "reverse the axes".
    {% for name in {"permute", "permute!", "reshape", "reshape!"} %}
      def {{name.id}}(first : Int, *rest : Int)
        {{name.id}}([first.to_i32] + rest.map(&.to_i32).to_a)
      end
    {% end %}
  end
end
