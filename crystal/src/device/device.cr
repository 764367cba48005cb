require "./lib_ph_gpu"

module Phase
  # Raised by every block-taking method of a device array: arbitrary Crystal blocks cannot run
  # on the device path, and they are never silently run on the CPU instead.
  class DeviceBlockError < Exception
    def initialize(method : String)
      super("#{method}: arbitrary blocks cannot run on the device path. Call #to_host (or #to_narr " \
            "on the host array) first if a per-element Crystal block is really what you need.")
    end
  end

  # Process-wide state of the device path: library initialisation, status -> exception
  # translation, and the data-dependent error word.
  module Device
    @@initialised = false

    def self.init(device : Int32 = 0) : Nil
      check LibPhGpu.ph_init(device)
      @@initialised = true
    end

    def self.ensure_init : Nil
      init unless @@initialised
    end

    def self.shutdown : Nil
      LibPhGpu.ph_shutdown if @@initialised
      @@initialised = false
    end

    # CUDA / NCCL failures and a missing device surface as RuntimeError (there is no CPU fallback).
    def self.check(status : Int32) : Nil
      return if status == 0
      raise RuntimeError.new("libphgpu status #{status}: #{String.new(LibPhGpu.ph_last_error_string)}")
    end

    # Wait for every launched operator. A guaranteed raise point: the flag word comes back in the
    # same synchronisation (`ph_d2h_flags` with nothing to copy).
    def self.sync : Nil
      read_checked(Pointer(Void).null, Pointer(Void).null, LibC::SizeT.new(0))
    end

    # Internal wait that never raises (keeps a host temporary alive until it is copied).
    def self.wait : Nil
      check LibPhGpu.ph_sync
    end

    # Every synchronising READ is a raise point: `(a + b).to_host` raises OverflowError like
    # `a + b` does on the CPU path (`Int32#+`), without an explicit `raise_pending`.
    def self.read_checked(dst_host : Void*, src_dev : Void*, nbytes : LibC::SizeT) : Nil
      check LibPhGpu.ph_d2h_flags(dst_host, src_dev, nbytes, out flags)
      raise_for(flags)
    end

    def self.take_flags : UInt32
      check LibPhGpu.ph_take_arith_flags(out flags)
      flags
    end

    # Kernels accumulate data-dependent errors in one device flag word; this synchronises,
    # clears it and raises the class the CPU path would have raised at the offending element
    # (Int32#+ -> OverflowError, Int#// -> DivisionByZeroError, Enumerable#max on NaN -> ArgumentError).
    def self.raise_pending : Nil
      raise_for(take_flags)
    end

    def self.raise_for(flags : UInt32) : Nil
      raise DivisionByZeroError.new if flags & LibPhGpu::FLAG_DIV0 != 0
      raise OverflowError.new if flags & LibPhGpu::FLAG_OVERFLOW != 0
      raise ArgumentError.new("Overflow: Int::MIN // -1, or a negative integer exponent") if flags & LibPhGpu::FLAG_ARGUMENT != 0
      raise ArgumentError.new("Comparison of NaN failed") if flags & LibPhGpu::FLAG_NAN != 0
    end

    # Element type -> dtype code. Only Crystal's primitive numbers and Bool have a device
    # representation; anything else is a compile-time error.
    def self.dtype(t : T.class) : Int32 forall T
      {% if T == Float32 %}
        LibPhGpu::DType::F32.value
      {% elsif T == Float64 %}
        LibPhGpu::DType::F64.value
      {% elsif T == Int32 %}
        LibPhGpu::DType::I32.value
      {% elsif T == Int64 %}
        LibPhGpu::DType::I64.value
      {% elsif T == UInt8 || T == Bool %}
        LibPhGpu::DType::U8.value
      {% elsif T == Int8 %}
        LibPhGpu::DType::I8.value
      {% elsif T == Int16 %}
        LibPhGpu::DType::I16.value
      {% elsif T == UInt16 %}
        LibPhGpu::DType::U16.value
      {% elsif T == UInt32 %}
        LibPhGpu::DType::U32.value
      {% elsif T == UInt64 %}
        LibPhGpu::DType::U64.value
      {% else %}
        {% raise "#{T} has no device representation (primitive numbers and Bool only)" %}
      {% end %}
    end
  end

  # Ref-counted (by the GC) owner of one device allocation. `DeviceNArray#reshape` aliases it
  # like `NArray#reshape` aliases its Slice, and device views keep their source's buffer alive.
  class DeviceBuffer
    getter ptr : Void*
    getter bytesize : Int64
    @parent : DeviceBuffer? = nil
    @home : Void* = Pointer(Void).null

    def initialize(@bytesize : Int64)
      Device.ensure_init
      @ptr = Pointer(Void).null
      Device.check LibPhGpu.ph_alloc(LibC::SizeT.new({@bytesize, 1_i64}.max), pointerof(@ptr))
      @home = LibPhGpu.ph_stream # the pool block is released on the stream it was handed out on
    end

    # The stream the block was allocated on (its release is ordered there).
    def home_stream : Void*
      if parent = @parent
        parent.home_stream
      else
        @home
      end
    end

    # A byte range of another buffer (one slice of a batched `slices` copy). It keeps its parent
    # alive and never frees: the parent releases the whole allocation when the last range is gone.
    def initialize(parent : DeviceBuffer, byte_offset : Int64, @bytesize : Int64)
      @parent = parent
      @ptr = (parent.ptr.as(UInt8*) + byte_offset).as(Void*)
    end

    # Eager release; the finalizer is only the safety net.
    def free : Nil
      return if @ptr.null? || @parent
      LibPhGpu.ph_free_on(@ptr, @home)
      @ptr = Pointer(Void).null
    end

    def finalize
      free
    end
  end
end
