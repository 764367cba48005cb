require "./device_n_array"

module Phase
  # A CUDA stream of the library. `stream.use { ... }` makes every array operation of the block
  # launch on it; `stream.wait(other)` orders it behind what `other` (nil = the library's own
  # stream) has queued so far. Twin of `Phase::Stream` in include/ph_pipeline.hpp.
  class Stream
    getter handle : Void*

    def initialize
      Device.ensure_init
      @handle = Pointer(Void).null
      Device.check LibPhGpu.ph_stream_create(pointerof(@handle))
    end

    # Runs the block with this stream current and returns the block's value.
    def use(&)
      saved = LibPhGpu.ph_stream
      Device.check LibPhGpu.ph_set_stream(@handle)
      begin
        yield
      ensure
        LibPhGpu.ph_set_stream(saved)
      end
    end

    def wait(other : Stream? = nil) : Nil
      Device.check LibPhGpu.ph_stream_wait(@handle, other ? other.handle : Pointer(Void).null)
    end

    def synchronize : Nil
      Device.check LibPhGpu.ph_stream_sync(@handle)
    end

    def close : Nil
      return if @handle.null?
      LibPhGpu.ph_stream_destroy(@handle)
      @handle = Pointer(Void).null
    end

    def finalize
      close
    end

    # The library's own stream waits for everything queued so far on `s`.
    def self.main_wait(s : Stream) : Nil
      Device.check LibPhGpu.ph_stream_wait(Pointer(Void).null, s.handle)
    end
  end

  # A row-major host array in PINNED memory (`ph_host_alloc`): the only kind of host memory an
  # asynchronous transfer may use (the Boehm heap behind `NArray`'s `Slice(T)` is pageable, and
  # the GC may free it while a copy is still queued).
  class PinnedArray(T)
    getter shape : Array(Int32)
    getter ptr : Pointer(T)

    def initialize(shape : Enumerable(Int))
      @shape = shape.map(&.to_i32).to_a
      raw = Pointer(Void).null
      Device.check LibPhGpu.ph_host_alloc(LibC::SizeT.new({size, 1_i64}.max * sizeof(T)), pointerof(raw))
      @ptr = raw.as(Pointer(T))
    end

    # A pinned copy of a host array.
    def self.from(src : NArray(T)) : self
      result = new(src.shape)
      src.buffer.copy_to(result.ptr, src.buffer.size)
      result
    end

    def size : Int64
      Descriptor.element_count(@shape)
    end

    def row_elems : Int64
      (@shape.empty? || @shape[0] == 0) ? 0_i64 : size // @shape[0]
    end

    # Address of row `r0` of the leading axis (no copy).
    def rows(r0 : Int) : Pointer(T)
      @ptr + r0.to_i64 * row_elems
    end

    def rows_shape(r0 : Int, r1 : Int) : Array(Int32)
      s = @shape.clone
      s[0] = (r1 - r0).to_i32
      s
    end

    # The contents as an ordinary `NArray` (a copy onto the GC heap).
    def to_narr : NArray(T)
      NArray.of_buffer(@shape.clone, Slice(T).new(size.to_i32) { |i| @ptr[i] })
    end

    def free : Nil
      return if @ptr.null?
      LibPhGpu.ph_host_free(@ptr.as(Void*))
      @ptr = Pointer(T).null
    end

    def finalize
      free
    end
  end

  class DeviceNArray(T)
    # Host -> device from pinned memory: returns at once, the copy is ordered on the current
    # stream like an operator.
    def self.from_host_async(shape : Enumerable(Int), pinned : Pointer(T)) : self
      result = new(shape)
      if result.size > 0
        Device.check LibPhGpu.ph_h2d(result.dev.ptr, pinned.as(Void*), LibC::SizeT.new(result.size * sizeof(T)))
      end
      result
    end

    # Device -> host into pinned memory, asynchronous: valid after the next `Device.sync` (the
    # raise point for data-dependent errors) or `Stream#synchronize`. `self` may be a temporary:
    # its block is released on the stream it was allocated on, which is made to wait for this
    # copy when that is not the current stream.
    def to_host_async(pinned_dst : Pointer(T)) : Nil
      if size > 0
        Device.check LibPhGpu.ph_d2h_async(pinned_dst.as(Void*), dev.ptr, LibC::SizeT.new(size * sizeof(T)))
      end
      home = dev.home_stream
      cur = LibPhGpu.ph_stream
      Device.check LibPhGpu.ph_stream_wait(home, cur) unless home == cur
    end
  end

  # Chunked host -> device -> host pipeline over the leading axis, written with the array API.
  # ONE elementwise expression over host-resident operands is bound by the host link, not by any
  # kernel; three streams by ROLE (upload, compute, download) let the upload of chunk i+1, the
  # kernels of chunk i and the download of chunk i-1 overlap. Twin of `Phase::RowPipeline` in
  # include/ph_pipeline.hpp and of ph-core_b200/pipeline.py.
  class RowPipeline
    # `chunks` equal chunks; with `taper` = t the LAST one is cut again into halves t times.
    # What is left when the last upload ends is one chunk's kernels and download -- nothing
    # overlaps that tail -- so the final chunks are small while the early ones stay large (every
    # copy pays a fixed set-up).
    def self.row_chunks(n : Int64, chunks : Int64, taper : Int32 = 0, ramp : Int32 = 0) : Array({Int64, Int64})
      per = (n + {chunks, 1_i64}.max - 1) // {chunks, 1_i64}.max
      bounds = [] of {Int64, Int64}
      r = 0_i64
      while r < n
        bounds << {r, {n, r + per}.min}
        r += per
      end
      if taper > 0 && !bounds.empty?
        r0, r1 = bounds.pop
        taper.times do
          mid = r0 + (r1 - r0 + 1) // 2
          break if mid >= r1
          bounds << {r0, mid}
          r0 = mid
        end
        bounds << {r0, r1}
      end
      # `ramp` = t: the mirror image at the FRONT (per/2^t, per/2^t, ..., per/2) -- the first download
      # can only start after the first chunk, so that one is small too.
      if ramp > 0 && !bounds.empty?
        r0, r1 = bounds.shift
        head = [] of {Int64, Int64}
        ramp.times do
          mid = r1 - (r1 - r0 + 1) // 2
          break if mid <= r0
          head.unshift({mid, r1})
          r1 = mid
        end
        head.unshift({r0, r1})
        bounds = head + bounds
      end
      bounds
    end

    def initialize(@chunks : Int64 = 4_i64, @taper : Int32 = 7, @ramp : Int32 = 5)
      @up = Stream.new
      @comp = Stream.new
      @down = Stream.new
    end

    # `dest[r0...r1] = yield(rows[k][r0...r1] as device arrays, shared as device arrays)` for every
    # row chunk. The block composes device operators (it runs on the host and only launches
    # kernels -- it is not a per-element block). `wait: true` returns when `dest` is complete and
    # raises pending data-dependent errors.
    def map_rows(rows : Array(PinnedArray(T)), dest : PinnedArray(T), shared : Array(PinnedArray(T)) = [] of PinnedArray(T),
                 wait : Bool = true, &block : Array(DeviceNArray(T)), Array(DeviceNArray(T)) -> DeviceNArray(T)) : Nil forall T
      n = dest.shape.empty? ? 0_i64 : dest.shape[0].to_i64
      rows.each do |r|
        if r.shape.empty? || r.shape[0] != n
          raise ShapeError.new("map_rows: every row operand needs the leading extent of the output")
        end
      end
      up, comp, down = @up, @comp, @down
      up.wait; comp.wait; down.wait # behind whatever the main stream has queued
      shared_dev = [] of DeviceNArray(T)
      up.use { shared.each { |x| shared_dev << DeviceNArray(T).from_host_async(x.shape, x.ptr) } }
      keep = [] of DeviceNArray(T)
      begin
        RowPipeline.row_chunks(n, @chunks, @taper, @ramp).each do |(r0, r1)|
          ins = [] of DeviceNArray(T)
          up.use { rows.each { |r| ins << DeviceNArray(T).from_host_async(r.rows_shape(r0, r1), r.rows(r0)) } }
          comp.wait(up) # chunk k's operands (and the shared ones) have landed
          res = comp.use { block.call(ins, shared_dev) }
          down.wait(comp)
          down.use { res.to_host_async(dest.rows(r0)) }
          keep.concat(ins)
          keep << res
        end
      rescue ex
        # the block raised half way: chunks already queued still use the temporaries, so nothing is
        # released (by the GC's finalizers) before the three streams have drained
        up.synchronize; comp.synchronize; down.synchronize
        raise ex
      end
      # device temporaries are released on the streams they were allocated on: order each of
      # those behind every consumer before letting go, then join the main stream
      up.wait(comp); up.wait(down); comp.wait(down)
      Stream.main_wait(up); Stream.main_wait(comp); Stream.main_wait(down)
      keep.each(&.dev.free)
      shared_dev.each(&.dev.free)
      Device.sync if wait
    end

    def close : Nil
      @up.close; @comp.close; @down.close
    end
  end
end
