-- Propagator configuration in the normal form serde_dhall writes (layout of the reference's data/02_config/prop_config.dhall;
-- values are this repository's test fixture: JGM-3 12x12 from the COF fixture, Earth + Moon point masses, exponential drag).
{ accel_models =
  { gravity_field = Some
    { _1 =
      { degree = 12
      , filepath = "jgm3_12x12.sha.tab"
      , gunzipped = False
      , order = 10
      }
    , _2 = { ephemeris_id = +399, orientation_id = +399 }
    }
  , point_masses = Some
    { celestial_objects = [ +301 ]
    , correction =
        None { converged : Bool, stellar : Bool, transmit_mode : Bool }
    }
  }
, force_models =
  { drag = Some
    { density =
        < Constant : Double
        | Exponential : { r0 : Double, ref_alt_m : Double, rho0 : Double }
        | StdAtm : { max_alt_m : Double }
        >.Exponential
          { r0 = 700000.0, ref_alt_m = 88667.0, rho0 = 3.614e-13 }
    , drag_frame = { ephemeris_id = +399, mu_km3_s2 = Some 398600.435436096, orientation_id = +399
      , shape = None { polar_radius_km : Double, semi_major_equatorial_radius_km : Double, semi_minor_equatorial_radius_km : Double } }
    , estimate = False
    }
  , solar_pressure =
      None
        { estimate : Bool
        , phi : Double
        , shadow_model :
            { light_source : { ephemeris_id : Integer, mu_km3_s2 : Optional Double, orientation_id : Integer }
            , shadow_bodies : List { ephemeris_id : Integer, mu_km3_s2 : Optional Double, orientation_id : Integer }
            }
        }
  }
, method =
    < CashKarp45 | DormandPrince45 | DormandPrince78 | RungeKutta4 | RungeKutta89 | Verner56 >.DormandPrince78
, options =
  { attempts = 30
  , error_ctrl =
      < LargestError | LargestState | LargestStep | RSSCartesianState | RSSCartesianStep | RSSState | RSSStep >.RSSCartesianState
  , fixed_step = False
  , init_step = "30 s"
  , integration_frame = None { ephemeris_id : Integer, orientation_id : Integer }
  , max_step = "10 min"
  , min_step = "1 ms"
  , tolerance = 1.0e-11
  }
}
