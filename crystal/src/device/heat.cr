require "./device_n_array"

module Phase
  # The stencil of examples/heat_equation.cr as one fused kernel per step (or per TWO steps on
  # 2-D / 3-D grids: temporal blocking, bit-identical to single steps).
  module Heat
    # `update_temp`: one explicit step. `Example1D` reproduces the example's one-sided ends
    # (`d[0] = (s[1] - s[0]) * C`); `Fixed` is the N-D rule written with the reference's operators:
    # `c = s[1...-1, ...]; d_k = (s[lo_k] - 2 * c) + s[hi_k]; nxt[1...-1, ...] = c + ((d_0 + d_1) + d_2) * C`
    # -- the same association order, every operation rounded on its own, no FMA.
    def self.update_temp(state : DeviceNArray(T), coeff : T, mode = LibPhGpu::HeatMode::Fixed) : DeviceNArray(T) forall T
      result = DeviceNArray(T).new(state.shape)
      if state.size > 0
        extents = state.shape.map(&.to_i64)
        Device.check LibPhGpu.ph_heat_step(Device.dtype(T), extents.size, extents.to_unsafe, pointerof(coeff).as(Void*), mode.value, state.dev.ptr, result.dev.ptr)
      end
      result
    end

    # `simulate`: `steps` steps; the initial state is left untouched.
    def self.simulate(initial : DeviceNArray(T), coeff : T, steps : Int, mode = LibPhGpu::HeatMode::Fixed) : DeviceNArray(T) forall T
      a = initial.clone
      return a if initial.size == 0 || steps <= 0
      b = DeviceNArray(T).new(initial.shape)
      extents = initial.shape.map(&.to_i64)
      final_is_b = 0
      Device.check LibPhGpu.ph_heat_run(Device.dtype(T), extents.size, extents.to_unsafe, pointerof(coeff).as(Void*), mode.value, a.dev.ptr, b.dev.ptr, steps.to_i64, pointerof(final_is_b))
      final_is_b != 0 ? b : a
    end
  end
end
