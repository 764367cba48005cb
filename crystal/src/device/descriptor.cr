require "./lib_ph_gpu"

module Phase
  # Descriptor compilation: the part of `IndexRegion`, `CoordUtil`, `ShapeUtil` and the View
  # transforms that the device path needs as ONE stride/offset record instead of a per-element
  # coordinate walk. Same rules as ph-core_b200/csrc/host_index.cpp (checked there against the
  # reference's own coordinate enumeration on random regions and transform chains).
  module Descriptor
    alias Desc = LibPhGpu::Desc

    # `Buffered.axis_strides` as a descriptor of a whole row-major array.
    def self.contiguous(shape : Indexable(Int)) : Desc
      raise DimensionError.new("The device path supports at most #{LibPhGpu::MAX_RANK} dimensions.") if shape.size > LibPhGpu::MAX_RANK
      d = Desc.new
      d.rank = shape.size
      acc = 1_i64
      (shape.size - 1).downto(0) do |i|
        d.extent[i] = shape[i].to_i64
        d.stride[i] = acc
        acc *= shape[i]
      end
      d
    end

    # `ShapeUtil.shape_to_size` in Int64 (the empty shape [] has size 0).
    def self.element_count(shape : Indexable(Int)) : Int64
      shape.empty? ? 0_i64 : shape.reduce(1_i64) { |acc, n| acc * n }
    end

    def self.count(d : Desc) : Int64
      n = 1_i64
      d.rank.times { |i| n *= d.extent[i] }
      n
    end

    # Replaces `IndexRegion#local_to_absolute_unsafe` + `Buffered.coord_to_index_fast` for a
    # whole region: offset += sum first*stride; kept axes get extent = proper_shape,
    # stride = step*stride; if every axis is dropped the result is the 1-element vector [size].
    def self.region(src : Desc, region : IndexRegion) : Desc
      d = Desc.new
      d.offset = src.offset
      first, step = region.first, region.stride
      unless step.any?(&.zero?) # an empty region's first/last are meaningless
        first.each_with_index { |f, i| d.offset += f.to_i64 * src.stride[i] }
      end
      rank = 0
      region.proper_shape.each_with_index do |n, i|
        next if region.drop && region.degeneracy[i]
        d.extent[rank] = n.to_i64
        d.stride[rank] = step[i].to_i64 * src.stride[i]
        rank += 1
      end
      if rank == 0
        d.extent[0] = region.shape[0].to_i64
        d.stride[0] = 1_i64
        rank = 1
      end
      d.rank = rank
      d
    end

    # `PermuteTransform`: output axis i is source axis pattern[i]; no pattern = reversed axes.
    def self.permute(src : Desc, pattern : Indexable(Int)? = nil) : Desc
      d = src
      n = pattern ? pattern.size : src.rank
      n.times do |i|
        from = pattern ? pattern[i].to_i32 : src.rank - 1 - i
        unless 0 <= from < src.rank
          raise IndexError.new("Could not use pattern to permute: axis #{from} is not present in a #{src.rank}-dimensional MultiIndexable.")
        end
        d.extent[i] = src.extent[from]
        d.stride[i] = src.stride[from]
      end
      d.rank = n
      d
    end

    # `ReverseTransform`: every axis flipped.
    def self.reverse(src : Desc) : Desc
      d = src
      d.rank.times do |i|
        d.offset += (d.extent[i] - 1) * d.stride[i] if d.extent[i] > 0
        d.stride[i] = -d.stride[i]
      end
      d
    end

    # `ReshapeTransform` when it is expressible in strides: the source splits into maximal runs
    # that are contiguous in lexicographic order and every new axis must subdivide one run.
    # Returns nil when a copy is needed first (the caller materialises and retries).
    def self.reshape(src : Desc, new_shape : Indexable(Int)) : Desc?
      old_n = src.rank == 0 ? 0_i64 : count(src)
      new_n = new_shape.empty? ? 0_i64 : new_shape.reduce(1_i64) { |acc, n| acc * n }
      if old_n != new_n
        raise ShapeError.new("Cannot change shape (#{old_n} elements) to #{new_shape.to_a} (#{new_n} elements) because reshape cannot add or remove elements.")
      end
      d = Desc.new
      d.rank = new_shape.size
      d.offset = src.offset
      new_shape.each_with_index { |n, i| d.extent[i] = n.to_i64 }
      return d if old_n == 0

      old_ext = [] of Int64
      old_str = [] of Int64
      src.rank.times do |i|
        next if src.extent[i] == 1
        old_ext << src.extent[i]
        old_str << src.stride[i]
      end
      new_idx = (0...new_shape.size).select { |i| new_shape[i] != 1 }

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
        stride = old_str[oj - 1]
        (nj - 1).downto(ni) do |k|
          d.stride[new_idx[k]] = stride
          stride *= new_shape[new_idx[k]]
        end
        oi, ni = oj, nj
      end
      d
    end

    # A broadcast operand is its own descriptor with stride 0 on every stretched axis.
    def self.broadcast(src : Desc, shape : Indexable(Int)) : Desc
      raise ShapeError.new("Broadcasting requires equal rank (#{src.rank} vs #{shape.size}).") if shape.size != src.rank
      d = src
      shape.each_with_index do |n, i|
        next if src.extent[i] == n
        raise ShapeError.new("Axis #{i} of length #{src.extent[i]} cannot be stretched to #{n}.") if src.extent[i] != 1
        d.extent[i] = n.to_i64
        d.stride[i] = 0_i64
      end
      d
    end

    # Buffer offset of one canonical coordinate (`Buffered.coord_to_index_fast`).
    def self.offset_of(d : Desc, coord : Indexable(Int)) : Int64
      off = d.offset
      coord.each_with_index { |c, i| off += c.to_i64 * d.stride[i] }
      off
    end
  end

  module ShapeUtil
    # NEW: the reference has no broadcasting (binary operators demand identical shapes). The
    # device path defines it as `tile` + operator: equal rank, each axis equal or 1.
    def self.broadcast_shapes(a : Indexable(Int), b : Indexable(Int)) : Array(Int32)
      raise ShapeError.new("Broadcasting requires equal rank (#{a.size} vs #{b.size}).") if a.size != b.size
      a.to_a.zip(b.to_a).map_with_index do |(x, y), i|
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
