require "./lib_ph_gpu"

module Phase
  # Descriptor compilation: the part of `IndexRegion`, `CoordUtil`, `ShapeUtil` and the View
  # transforms that the device path needs as ONE stride/offset record instead of a per-element
  # coordinate walk. Same rules as ph-core_b200/csrc/host_index.cpp (checked there against the
  # reference's own coordinate enumeration on random regions and transform chains).
  #
  # NB `LibPhGpu::Desc#extent` / `#stride` are StaticArrays, i.e. VALUES: `d.extent[i] = x` would
  # write into a temporary copy. Every function below fills local arrays and assigns them whole.
  module Descriptor
    alias Desc = LibPhGpu::Desc
    alias Axes = StaticArray(Int64, 8)

    def self.axes : Axes
      Axes.new(0_i64)
    end

    def self.make(rank : Int, offset : Int64, extent : Axes, stride : Axes) : Desc
      d = Desc.new
      d.rank = rank.to_i32
      d.offset = offset
      d.extent = extent
      d.stride = stride
      d
    end

    private def self.check_rank(rank : Int)
      if rank > LibPhGpu::MAX_RANK
        raise DimensionError.new("The device path supports at most #{LibPhGpu::MAX_RANK} dimensions (got #{rank}).")
      end
    end

    # `Buffered.axis_strides` as a descriptor of a whole row-major array.
    def self.contiguous(shape : Indexable(Int)) : Desc
      check_rank(shape.size)
      extent, stride = axes, axes
      acc = 1_i64
      (shape.size - 1).downto(0) do |i|
        extent[i] = shape[i].to_i64
        stride[i] = acc
        acc *= shape[i]
      end
      make(shape.size, 0_i64, extent, stride)
    end

    # The slice `src[.., index, ..]` along `axis` as a descriptor over the same buffer: the axis is removed,
    # the offset moves by index * stride (what `each_slice` hands out; a vector's slices keep shape [1]).
    def self.drop_axis(src : Desc, axis : Int, index : Int64) : Desc
      extent, stride = axes, axes
      src_extent, src_stride = src.extent, src.stride
      rank = 0
      src.rank.times do |i|
        next if i == axis
        extent[rank] = src_extent[i]
        stride[rank] = src_stride[i]
        rank += 1
      end
      if rank == 0
        extent[0] = 1_i64
        stride[0] = 1_i64
        rank = 1
      end
      make(rank, src.offset + index * src_stride[axis], extent, stride)
    end

    # `ShapeUtil.shape_to_size` in Int64 (the empty shape [] has size 0).
    def self.element_count(shape : Indexable(Int)) : Int64
      shape.empty? ? 0_i64 : shape.reduce(1_i64) { |acc, n| acc * n }
    end

    def self.count(d : Desc) : Int64
      extent = d.extent
      n = 1_i64
      d.rank.times { |i| n *= extent[i] }
      n
    end

    # Replaces `IndexRegion#local_to_absolute_unsafe` + `Buffered.coord_to_index_fast` for a
    # whole region: offset += sum first*stride; kept axes get extent = proper_shape,
    # stride = step*stride; if every axis is dropped the result is the 1-element vector [size].
    def self.region(src : Desc, region : IndexRegion) : Desc
      src_stride = src.stride
      offset = src.offset
      first, step = region.first, region.stride
      unless step.any?(&.zero?) # an empty region's first/last are meaningless
        first.each_with_index { |f, i| offset += f.to_i64 * src_stride[i] }
      end
      extent, stride = axes, axes
      rank = 0
      region.proper_shape.each_with_index do |n, i|
        next if region.drop && region.degeneracy[i]
        extent[rank] = n.to_i64
        stride[rank] = step[i].to_i64 * src_stride[i]
        rank += 1
      end
      if rank == 0
        extent[0] = region.shape[0].to_i64
        stride[0] = 1_i64
        rank = 1
      end
      make(rank, offset, extent, stride)
    end

    # `PermuteTransform`: output axis i is source axis pattern[i]; no pattern = reversed axes.
    def self.permute(src : Desc, pattern : Indexable(Int)? = nil) : Desc
      src_extent, src_stride = src.extent, src.stride
      n = pattern ? pattern.size : src.rank
      check_rank(n)
      extent, stride = axes, axes
      n.times do |i|
        from = pattern ? pattern[i].to_i32 : src.rank - 1 - i
        unless 0 <= from < src.rank
          raise IndexError.new("Could not use pattern to permute: axis #{from} is not present in a #{src.rank}-dimensional MultiIndexable.")
        end
        extent[i] = src_extent[from]
        stride[i] = src_stride[from]
      end
      make(n, src.offset, extent, stride)
    end

    # `ReverseTransform`: every axis flipped.
    def self.reverse(src : Desc) : Desc
      extent, stride = src.extent, src.stride
      offset = src.offset
      src.rank.times do |i|
        offset += (extent[i] - 1) * stride[i] if extent[i] > 0
        stride[i] = -stride[i]
      end
      make(src.rank, offset, extent, stride)
    end

    # `ReshapeTransform` when it is expressible in strides: the source splits into maximal runs
    # that are contiguous in lexicographic order and every new axis must subdivide one run.
    # Returns nil when a copy is needed first (the caller materialises and retries).
    def self.reshape(src : Desc, new_shape : Indexable(Int)) : Desc?
      check_rank(new_shape.size)
      src_extent, src_stride = src.extent, src.stride
      old_n = src.rank == 0 ? 0_i64 : count(src)
      new_n = element_count(new_shape)
      if old_n != new_n
        raise ShapeError.new("Cannot change shape (#{old_n} elements) to #{new_shape.to_a} (#{new_n} elements) because reshape cannot add or remove elements.")
      end
      extent, stride = axes, axes
      new_shape.each_with_index { |n, i| extent[i] = n.to_i64 }
      return make(new_shape.size, src.offset, extent, stride) if old_n == 0

      old_ext = [] of Int64
      old_str = [] of Int64
      src.rank.times do |i|
        next if src_extent[i] == 1
        old_ext << src_extent[i]
        old_str << src_stride[i]
      end
      new_idx = (0...new_shape.size).select { |i| new_shape[i] != 1 }.to_a

      oi = ni = 0
      while oi < old_ext.size && ni < new_idx.size
        oj, nj = oi + 1, ni + 1
        op, np = old_ext[oi], new_shape[new_idx[ni]].to_i64
        while op != np
          if op < np
            op *= old_ext[oj]
            oj += 1
          else
            np *= new_shape[new_idx[nj]]
            nj += 1
          end
        end
        (oi...oj - 1).each do |k|
          return nil if old_str[k] != old_str[k + 1] * old_ext[k + 1]
        end
        run = old_str[oj - 1]
        (nj - 1).downto(ni) do |k|
          stride[new_idx[k]] = run
          run *= new_shape[new_idx[k]]
        end
        oi, ni = oj, nj
      end
      make(new_shape.size, src.offset, extent, stride)
    end

    # A broadcast operand is its own descriptor with stride 0 on every stretched axis.
    def self.broadcast(src : Desc, shape : Indexable(Int)) : Desc
      raise ShapeError.new("Broadcasting requires equal rank (#{src.rank} vs #{shape.size}).") if shape.size != src.rank
      extent, stride = src.extent, src.stride
      shape.each_with_index do |n, i|
        next if extent[i] == n
        raise ShapeError.new("Axis #{i} of length #{extent[i]} cannot be stretched to #{n}.") if extent[i] != 1
        extent[i] = n.to_i64
        stride[i] = 0_i64
      end
      make(src.rank, src.offset, extent, stride)
    end

    # `MultiIndexable#tile` as a gather: every axis becomes (count, extent) with strides (0, stride),
    # so the device never computes a modulo. Returns {source, contiguous destination}.
    def self.tile(src : Desc, counts : Indexable(Int)) : {Desc, Desc}
      check_rank(2 * counts.size)
      src_extent, src_stride = src.extent, src.stride
      extent, stride = axes, axes
      doubled = [] of Int64
      counts.each_with_index do |c, i|
        extent[2 * i] = c.to_i64
        stride[2 * i] = 0_i64
        extent[2 * i + 1] = src_extent[i]
        stride[2 * i + 1] = src_stride[i]
        doubled << c.to_i64 << src_extent[i]
      end
      {make(2 * counts.size, src.offset, extent, stride), contiguous(doubled)}
    end

    # Buffer offset of one canonical coordinate (`Buffered.coord_to_index_fast`).
    def self.offset_of(d : Desc, coord : Indexable(Int)) : Int64
      stride = d.stride
      off = d.offset
      coord.each_with_index { |c, i| off += c.to_i64 * stride[i] }
      off
    end
  end

  module ShapeUtil
    # NEW: the reference has no broadcasting (binary operators demand identical shapes). The
    # device path defines it as `tile` + operator: equal rank, each axis equal or 1.
    def self.broadcast_shapes(a : Indexable(Int), b : Indexable(Int)) : Array(Int32)
      raise ShapeError.new("Broadcasting requires equal rank (#{a.size} vs #{b.size}).") if a.size != b.size
      Array(Int32).new(a.size) do |i|
        x, y = a[i], b[i]
        if x == y || y == 1
          x.to_i32
        elsif x == 1
          y.to_i32
        else
          raise ShapeError.new("Shapes #{a.to_a} and #{b.to_a} cannot be broadcast on axis #{i}.")
        end
      end
    end
  end

  struct IndexRegion(T)
    # The one field the descriptor compile needs beyond `#first` / `#stride` / `#shape` /
    # `#degeneracy` / `#drop`; upstream keeps it as an ivar without a getter.
    def proper_shape : Array(T)
      @proper_shape
    end

    def to_descriptor(src : LibPhGpu::Desc) : LibPhGpu::Desc
      Descriptor.region(src, self)
    end
  end
end
