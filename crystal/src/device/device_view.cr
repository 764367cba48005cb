require "./device_indexable"

module Phase
  # Lazy view of a device array: source buffer + ONE descriptor. `View` composes
  # Region / Permute / Reverse / Reshape coordinate transforms and applies the chain per element
  # (two heap allocations each); on the device the first three are affine in the coordinate, so
  # the chain folds into (offset, extent[], stride[]) as it is built and a read is one gather, a
  # write one scatter. Reshape folds when strides can express it and materialises otherwise.
  # It is both `View` and `MutableView`: writes go through to the source.
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

    def clone : self
      DeviceView(T).new(@dev, @desc, @shape.clone)
    end

    def view(region = nil) : DeviceView(T)
      new_view = clone
      new_view.restrict_to(region) if region
      new_view
    end

    def mutable_view(region = nil) : DeviceView(T)
      view(region)
    end

    protected def restrict_to(region : Enumerable) : self
      restrict_to(IndexRegion.new(region, @shape))
    end

    protected def restrict_to(region : IndexRegion) : self
      @desc = region.to_descriptor(@desc)
      @shape = region.shape
      self
    end

    # A chunk of a view is a view, like `View#unsafe_fetch_chunk`.
    def unsafe_fetch_chunk(region : IndexRegion) : DeviceView(T)
      view(region)
    end

    def reshape!(new_shape) : self
      new_shape = new_shape.map(&.to_i32).to_a
      if ShapeUtil.shape_to_size(new_shape) != size
        raise ShapeError.new("Cannot change shape from #{@shape.join('x')} (#{size} elements) to #{new_shape.join('x')} (#{ShapeUtil.shape_to_size(new_shape)} elements) because reshape cannot add or remove elements.")
      end
      if folded = Descriptor.reshape(@desc, new_shape)
        @desc = folded
      else
        copy = to_narr # not expressible in strides: materialise, then reshape the copy
        @dev = copy.dev
        @desc = Descriptor.contiguous(new_shape)
      end
      @shape = new_shape
      self
    end

    def reshape(new_shape) : self
      clone.reshape!(new_shape)
    end

    def permute!(order : Enumerable? = nil) : self
      if order && (bad_axis = order.find { |axis| axis < 0 || axis >= @shape.size })
        raise IndexError.new("Could not use pattern #{order} to permute: Axis #{bad_axis} is not present in a #{dimensions}-dimensional MultiIndexable")
      end
      @desc = Descriptor.permute(@desc, order.try &.to_a)
      extent = @desc.extent
      @shape = Array(Int32).new(@desc.rank) { |i| extent[i].to_i32 }
      self
    end

    def permute(order : Enumerable? = nil) : self
      clone.permute!(order)
    end

    # Splat forms (`view.permute(1, 0, 2)`); at least one argument, so that a bare `permute`
    # always means "reverse the axes".
    {% for name in {"permute", "permute!", "reshape", "reshape!"} %}
      def {{name.id}}(first : Int, *rest : Int)
        {{name.id}}([first.to_i32] + rest.map(&.to_i32).to_a)
      end
    {% end %}

    def reverse! : self
      @desc = Descriptor.reverse(@desc)
      self
    end

    def reverse : self
      clone.reverse!
    end
  end
end
