(ns raytrace-clj.dump-vectors
  "Golden vectors from the REAL reference, for whoever has a JVM (the build environment of libraytrace_b200 has none).

  Put this file at src/raytrace_clj/dump_vectors.clj of an UNMODIFIED gonewest818/raytrace-clj checkout and run

      lein run -m raytrace-clj.dump-vectors > jvm_vectors.json

  then copy jvm_vectors.json to tests/golden/jvm_vectors.json of the raytrace-clj_b200 repository and run
  `python -m pytest tests/test_jvm_vectors.py`: the CPU oracle (oracle/oracle.cpp, the restatement every GPU
  parity test is checked against) is compared with these numbers to 1e-12 relative.

  What is dumped (everything the reference's own tests do NOT pin numerically, SURVEY 8c):
    hit? of Sphere / UVSphere / MovingSphere / RectXY/XZ/YZ / Triangle / FlipNormals / Translate / RotateY / Box /
    AABB (t, p, normal, uv), scatter + emitted of every Shader with the random draws replaced by fixed values
    (with-redefs on rand-in-unit-sphere and rand), sample of every Texture that needs no file, get-ray of both
    cameras, and `pixel`'s gamma / quantisation on fixed colours."
  (:require [clojure.core.matrix :as mat]
            [clojure.string :as str]
            [raytrace-clj.util :as u :refer [vec3 ray]]
            [raytrace-clj.hitable :as hit]
            [raytrace-clj.shader :as shad]
            [raytrace-clj.texture :as tex]
            [raytrace-clj.camera :as cam]))

(mat/set-current-implementation :vectorz)

(defn- v [x] (if (nil? x) nil (mapv double (seq x))))
(defn- json [x]
  (cond (nil? x) "null"
        (map? x) (str "{" (str/join "," (for [[k val] x] (str "\"" (name k) "\":" (json val)))) "}")
        (sequential? x) (str "[" (str/join "," (map json x)) "]")
        (string? x) (str "\"" x "\"")
        (keyword? x) (str "\"" (name x) "\"")
        (true? x) "true" (false? x) "false"
        (and (number? x) (Double/isNaN (double x))) "\"nan\""
        (and (number? x) (Double/isInfinite (double x))) (if (pos? x) "\"inf\"" "\"-inf\"")
        (integer? x) (str x)
        :else (let [d (double x)] (.toString (java.math.BigDecimal. d)))))   ; exact decimal expansion of the double

(defn- hrec->map [h] (when h {:t (:t h) :p (v (:p h)) :normal (v (:normal h)) :uv (v (:uv h))}))

(def rays
  (vec (for [[o d] [[[13 2 3] [-13 -1 -3]] [[13 2 3] [-12.5 -1.7 -2.4]] [[0 5 0] [0.1 -1 0.05]] [[0.5 0.5 -10] [0 0 2]]
                    [[278 278 -800] [0.1 -0.2 1]] [[278 278 -800] [-0.3 0.4 1]] [[4 1.3 0.2] [-1 -0.2 0.1]]
                    [[0.3 0.2 0.1] [1 1 1]] [[-2 1 0] [1 0 0]] [[3 4 5] [-0.3 -0.4 -0.5]]]
             time [0.0 0.37]]
         {:o o :d d :time time})))
(defn- mk-ray [{:keys [o d time]}] (ray (apply vec3 o) (apply vec3 d) time))

(def gray (shad/lambertian :albedo (tex/constant :color (vec3 0.5 0.5 0.5))))

(def hitables
  {:sphere        [(hit/sphere :center (vec3 0 1 0) :radius 1 :material gray) {:center [0 1 0] :radius 1}]
   :ground        [(hit/sphere :center (vec3 0 -1000 0) :radius 1000 :material gray) {:center [0 -1000 0] :radius 1000}]
   :uv-sphere     [(hit/uv-sphere :center (vec3 0 0 0) :radius 1000 :material gray) {:center [0 0 0] :radius 1000}]
   :moving-sphere [(hit/moving-sphere :center0 (vec3 4 0.2 0) :t0 0.0 :center1 (vec3 4 0.6 0) :t1 1.0 :radius 0.2 :material gray)
                   {:center0 [4 0.2 0] :center1 [4 0.6 0] :t0 0 :t1 1 :radius 0.2}]
   :rect-xy       [(hit/rect-xy :x0 -1 :y0 -1 :x1 2 :y1 3 :k 0.5 :material gray) {:q [-1 -1 2 3 0.5]}]
   :rect-xz       [(hit/rect-xz :x0 0 :z0 0 :x1 555 :z1 555 :k 555 :material gray) {:q [0 0 555 555 555]}]
   :rect-yz       [(hit/rect-yz :y0 0 :z0 0 :y1 555 :z1 555 :k 0 :material gray) {:q [0 0 555 555 0]}]
   :flip-rect-xz  [(hit/flip-normals :item (hit/rect-xz :x0 0 :z0 0 :x1 555 :z1 555 :k 555 :material gray)) {:q [0 0 555 555 555]}]
   :triangle      [(hit/triangle :v0 (vec3 0 0 0) :v1 (vec3 0 1 0) :v2 (vec3 1 0 0) :material gray) {:q [0 0 0 0 1 0 1 0 0]}]
   :block         [(hit/translate :item (hit/rotate-y :item (hit/box :p0 (vec3 0 0 0) :p1 (vec3 165 330 165) :material gray)
                                                     :theta 15.0)
                                  :offset (vec3 265 0 295))
                   {:p0 [0 0 0] :p1 [165 330 165] :theta 15.0 :offset [265 0 295]}]})

(defn -main [& _]
  (let [fmax (double Float/MAX_VALUE)
        hits (for [[k [h params]] hitables, r rays, [tmin tmax] [[0.001 fmax] [0.0 7.5]]]
               {:kind k :params params :ray r :tmin tmin :tmax tmax
                :hit (hrec->map (hit/hit? h (mk-ray r) tmin tmax))
                :bbox (let [b (hit/bbox h 0.0 1.0)] {:vmin (v (:vmin b)) :vmax (v (:vmax b))})})
        aabbs (for [r rays, [lo hi] [[[-1 -1 -1] [1 1 1]] [[0 0 0] [555 555 555]]]]
                {:vmin lo :vmax hi :ray r
                 :hit (boolean (hit/hit? (hit/aabb :vmin (apply vec3 lo) :vmax (apply vec3 hi)) (mk-ray r) 0.001 fmax))})
        ball [0.1 -0.2 0.3]
        mats {:lambertian (shad/lambertian :albedo (tex/checkerboard :tex0 (tex/constant :color (vec3 0.2 0.3 0.1))
                                                                    :tex1 (tex/constant :color (vec3 0.9 0.9 0.9)) :scale 10))
              :metal (shad/metal :albedo (tex/constant :color (vec3 0.7 0.6 0.5)) :fuzz 0.3)
              :dielectric (shad/dielectric :ri 1.5)
              :light (shad/diffuse-light :tex (tex/uv-gradient :co (vec3 1 1 1) :cu (vec3 1 1 1) :cv (vec3 0.5 0.7 1.0) :cuv (vec3 0.5 0.7 1.0)))
              :isotropic (shad/isotropic :albedo (tex/constant :color (vec3 0.2 0.4 0.9)))}
        scat (for [[mk m] mats, r rays, rnd [0.01 0.5 0.99]
                   :let [s (hit/uv-sphere :center (vec3 0 1 0) :radius 1 :material m)
                         h (hit/hit? s (mk-ray r) 0.001 fmax)]
                   :when h]
               (with-redefs [u/rand-in-unit-sphere (fn [] (apply vec3 ball))
                             clojure.core/rand (fn ([] rnd) ([n] (* n rnd)))]
                 (let [sc (shad/scatter m (mk-ray r) h)]
                   {:material mk :ray r :ball ball :rand rnd :hit (hrec->map h)
                    :scattered (when sc {:o (v (:origin (:scattered sc))) :d (v (:direction (:scattered sc)))
                                         :time (:time (:scattered sc)) :attenuation (v (:attenuation sc))})
                    :emitted (v (shad/emitted m (:uv h) (:p h)))})))
        tl (cam/thin-lens-camera :lookfrom (vec3 13 2 3) :lookat (vec3 0 0 0) :vup (vec3 0 1 0) :vfov 20
                                 :aspect (/ (float 1200) (float 800)) :aperture 0.1 :focus-dist 10.0 :t0 0.0 :t1 1.0)
        ph (cam/pinhole-camera :lookfrom (vec3 13 2 3) :lookat (vec3 0 0 0) :vup (vec3 0 1 0) :vfov 20 :aspect 1.5)
        cams (for [[ck c] {:thin-lens tl :pinhole ph}, [s t] [[0.5 0.5] [0.0 1.0] [0.123 0.877]]]
               (with-redefs [u/rand-in-unit-disk (fn [] (vec3 0.3 -0.4 0)) clojure.core/rand (fn ([] 0.25) ([n] (* n 0.25)))]
                 (let [r (cam/get-ray c s t)]
                   {:camera ck :record (into {} (for [[k x] c] [k (if (number? x) x (v x))]))
                    :s s :t t :disk [0.3 -0.4] :rand 0.25 :o (v (:origin r)) :d (v (:direction r)) :time (:time r)})))
        gamma (for [c [[0.0 0.25 1.0] [0.5 2.0 7.0] [1e-6 0.999 0.1234]]]
                {:mean c :rgb8 (vec (seq (mat/emap #(int (min 255.99 %)) (mat/mul 255.99 (mat/sqrt (apply vec3 c))))))})]
    (println (json {:reference "gonewest818/raytrace-clj" :hits (vec hits) :aabb (vec aabbs) :scatter (vec scat)
                    :get_ray (vec cams) :gamma (vec gamma)}))))
